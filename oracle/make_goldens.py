"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/ref_harness.py).

Run from the repo root in the build container (the only place /root/reference exists):

    python oracle/make_goldens.py

The goldens pin the CPU oracle (oracle/*.c) -- and through it the CUDA path -- to the reference's
own outputs.  The reference ships no tests/fixtures, so these files are the only "known answers".
Versions are recorded inside every file.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL_ROOM = {  # 4 x 3 m -> 80 x 60 grid; two doors for one box => multi-door key
    "room_length": 4.0, "room_height": 3.0,
    "initial_boxes": {"b": [1.0, 1.5, 1.2, 1.6, 2.0, "d1", "d2"]},
    "targets": {"d1": [4.0, 1.5, 0.5, 0.8], "d2": [2.5, 3.0, 0.8, 0.5]},
    "walls": {"w": [2.0, 1.0, 0.2, 2.0]}, "holes": {"h": [2.0, 1.2, 0.3, 0.5]},
    "cylinders": {"c": [3.0, 1.6, 0.25]}}

DENSE_ROOM = {  # SURVEY App. B "dense" probe: 8 x 5 m, N = 30
    "room_length": 8.0, "room_height": 5.0,
    "initial_boxes": {"box": [2.0, 2.5, 3.0, 4.0, 2.5, "door_1"]},
    "targets": {"door_1": [8.0, 2.5, 0.6, 1.0]},
    "walls": {}, "holes": {}, "cylinders": {"cyl": [6.0, 2.5, 0.3]}}


def versions():
    return json.dumps({"numpy": np.__version__, "scipy": scipy.__version__,
                       "python": sys.version.split()[0]})


def hjb_golden(name, room, T, rooms=None, with_density=False, slices=(0, 1, 2, -2, -1)):
    m = None
    if with_density:  # a smooth synthetic density, 0 <= m <= ~1 (recompute=True input, simulations.py:435)
        with rh.RefEnv(rooms) as env:
            np.random.seed(0)
            with env.silence():
                simu = env.simulations.simulation(room, T)
                m = simu.gaussian_density(simu.sigma_convolution)
    out = rh.run_hjb(room, T, m=m, rooms=rooms)
    save = {"versions": versions(), "T": T, "room": json.dumps(rooms[room] if rooms and room in rooms else
                                                              json.load(open(os.path.join("rooms", room + ".json"))))}
    for kid, (key, o) in enumerate(out.items()):
        tr = o["trace"]
        nt = o["nt_opt"]
        sl = sorted({s % (nt - 1) for s in slices} | {(nt - 1) // 2})
        save.update({
            f"k{kid}_key": key, f"k{kid}_V": o["V"], f"k{kid}_nt": nt,
            f"k{kid}_nfev": tr["nfev"][0], f"k{kid}_h0": tr["h0"][0],
            f"k{kid}_attempt_h": np.array(tr["attempt_h"]), f"k{kid}_attempt_err": np.array(tr["attempt_err"]),
            f"k{kid}_slices": np.array(sl),
            f"k{kid}_vx": o["vx_opt"][sl], f"k{kid}_vy": o["vy_opt"][sl],
            # phi at the same slices: slice s <-> sol.y[:, nt-1-s]
            f"k{kid}_phi": np.stack([o["sol_y"][:, nt - 1 - s] for s in sl]),
            # per-slice checksums over ALL slices so every slice is pinned without storing it
            f"k{kid}_vx_sum": o["vx_opt"].sum(axis=(1, 2)), f"k{kid}_vy_sum": o["vy_opt"].sum(axis=(1, 2)),
            f"k{kid}_vx_abs": np.abs(o["vx_opt"]).sum(axis=(1, 2)),
        })
        if m is not None:
            save[f"k{kid}_m"] = m
    save["n_keys"] = len(out)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in save.items() if k.endswith("_vx")})


def gcfm_golden(name, room, T, rooms=None, n_steps=160, field_every=8, seed=0):
    """Free-running reference run; records everything needed to teacher-force any recorded step."""
    with rh.RefEnv(rooms) as env:
        np.random.seed(seed)
        with env.silence():
            simu = env.simulations.simulation(room, T)
        keys = list(simu.targets)
        with env.silence():
            for k in keys:
                simu.targets[k].compute_optimal_velocity(0.0, np.zeros((simu.Ny, simu.Nx)))
        rec = {"perm": [], "noise": [], "before": [], "after": [], "inside": []}
        real_choice, real_normal = np.random.choice, np.random.normal
        cur = {}

        def choice(a, size=None, replace=True, p=None):
            r = real_choice(a, size, replace, p)
            cur["perm"] = np.array(r)
            return r

        def normal(loc=0.0, scale=1.0, size=None):
            r = real_normal(loc, scale, size)
            cur.setdefault("noise", []).append(np.array(r))
            return r

        np.random.choice, np.random.normal = choice, normal
        try:
            fields = {}
            for s in range(n_steps):
                if simu.inside == 0:
                    break
                cur.clear()
                rec["before"].append(rh.agent_state(simu))
                simu.step(simu.dt)
                rec["after"].append(rh.agent_state(simu))
                rec["perm"].append(cur["perm"])
                nz = np.array(cur.get("noise", np.zeros((0, 2)))).reshape(-1, 2)
                pad = np.full((simu.N, 2), np.nan)
                pad[: len(nz)] = nz
                rec["noise"].append(pad)
                exited = (len(rec["inside"]) == 0 and simu.inside < simu.N) or \
                         (len(rec["inside"]) > 0 and simu.inside < rec["inside"][-1])
                rec["inside"].append(simu.inside)
                if s % field_every == 0 or exited:  # every step with an exit is teacher-forceable too
                    rec.setdefault("field_steps", []).append(s)
                    for kid, k in enumerate(keys):
                        fields[f"k{kid}_vx_step{s}"] = simu.targets[k].vx_opt[s].copy()
                        fields[f"k{kid}_vy_step{s}"] = simu.targets[k].vy_opt[s].copy()
        finally:
            np.random.choice, np.random.normal = real_choice, real_normal
        var_room = rooms[room] if rooms and room in rooms else json.load(open(os.path.join("rooms", room + ".json")))
        save = {"versions": versions(), "T": T, "room": json.dumps(var_room), "seed": seed,
                "field_every": field_every, "n_keys": len(keys),
                "v_des": np.array([a.v_des for a in simu.agents]),
                "agent_key": np.array([keys.index(a.target) for a in simu.agents]),
                "perm": np.array(rec["perm"]), "noise": np.array(rec["noise"]),
                "before": np.array(rec["before"]), "after": np.array(rec["after"]),
                "inside": np.array(rec["inside"]), "Vglobal": simu.V, "field_steps": np.array(rec["field_steps"]),
                "density_step_last": simu.gaussian_density(simu.sigma_convolution)}
        for kid, k in enumerate(keys):
            save[f"k{kid}_key"] = k
            save[f"k{kid}_V"] = simu.Vs[k]
            save[f"k{kid}_nt"] = simu.targets[k].nt_opt
        save.update(fields)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
        print("wrote", name, "steps", len(rec["perm"]), "N", simu.N, "inside", simu.inside)


def unit_golden(name="units"):
    """Direct calls of the reference's methods on random inputs (pair force, wall force, sampler,
    rasteriser, density)."""
    rooms = {"dense": DENSE_ROOM, "small": SMALL_ROOM}
    rng = np.random.RandomState(123)
    save = {"versions": versions()}
    with rh.RefEnv(rooms) as env:
        np.random.seed(1)
        with env.silence():
            simu = env.simulations.simulation("room_test", 1.0)
        a = simu.agents[0]
        key = list(simu.targets)[0]
        save["room"] = json.dumps(json.load(open("rooms/room_test.json")))
        # --- pair force (pedestrians.py:216-280)
        n = 3000
        pi_ = rng.uniform(1, 5, (n, 2)); pj = pi_ + rng.uniform(-3, 3, (n, 2))
        vi = rng.normal(0, 1.0, (n, 2)); vj = rng.normal(0, 1.0, (n, 2))
        vi[:50] = 0.0  # k = 0 branch (pedestrians.py:259-262)
        vj[25:75] = 0.0
        pj[100:110, 1] = pi_[100:110, 1]  # R_y = 0 -> atan2(-0.0, .) branch
        vd = rng.normal(1.34, 0.26, n)
        out = np.empty((n, 2))
        for q in range(n):
            a.traj[-1] = pi_[q]; a.vels[-1] = vi[q]; a.v_des = vd[q]
            out[q] = a.agents_repulsion(pj[q], vj[q])
        save.update(pair_pi=pi_, pair_pj=pj, pair_vi=vi, pair_vj=vj, pair_vdes=vd, pair_out=out)
        # --- wall force (pedestrians.py:282-334) on the room_test potential
        n = 400
        V = simu.Vs[key]
        pw = np.column_stack([rng.uniform(0.3, 9.7, n), rng.uniform(0.3, 5.7, n)])
        pw[:20] = np.round(pw[:20] / 0.05) * 0.05 + 0.025  # equidistant from nodes -> argmin ties
        vw = rng.normal(0, 1.0, (n, 2)); vw[:10] = 0
        vdw = rng.normal(1.34, 0.26, n)
        outw = np.empty((n, 2)); ind = np.empty(n, dtype=np.int64)
        for q in range(n):
            a.traj[-1] = pw[q]; a.vels[-1] = vw[q]; a.v_des = vdw[q]
            outw[q] = a.wall_repulsion(simu.X_opt, simu.Y_opt, V)
            d = np.sqrt((simu.X_opt - pw[q, 0]) ** 2 + (simu.Y_opt - pw[q, 1]) ** 2)
            ind[q] = np.argmin(d + V * 10e3)
        save.update(wall_V=V, wall_p=pw, wall_v=vw, wall_vdes=vdw, wall_out=outw, wall_ind=ind)
        # --- sampler (optimals.py:212-250) on a random field
        opt = simu.targets[key]
        opt.nt_opt = 5
        opt.vx_opt = rng.normal(size=(4, simu.Ny - 2, simu.Nx - 2)); opt.vy_opt = rng.normal(size=opt.vx_opt.shape)
        n = 600
        ps = np.column_stack([rng.uniform(0.0, 10.0, n), rng.uniform(0.0, 6.0, n)])
        ps[:40, 0] = rng.choice([0.0, 0.05, 0.049999, 0.050001, 0.1, 9.9, 9.84, 9.8, 0.15000000000000002], 40)
        ps[40:80, 1] = rng.choice([0.0, 0.05, 0.1, 5.9, 5.84, 5.8, 0.3, 0.35000000000000003], 40)
        ts = rng.randint(0, 6, n)
        outs = np.full((n, 2), np.nan); ok = np.zeros(n, dtype=np.uint8)
        for q in range(n):
            try:
                # the reference wraps negative indices silently; positions here are >= 0
                outs[q] = opt.choose_optimal_velocity(ps[q], int(ts[q])); ok[q] = 1
            except IndexError:
                ok[q] = 0
        save.update(samp_vx=opt.vx_opt, samp_vy=opt.vy_opt, samp_p=ps, samp_t=ts, samp_out=outs, samp_ok=ok,
                    samp_nt_opt=5)
        # --- rasteriser (simulations.py:516-576 + optimals.py:89-91) and density (simulations.py:453-487)
        for rname in ("room_test", "exit_opposite", "dense", "small"):
            np.random.seed(2)
            with env.silence():
                s2 = env.simulations.simulation(rname, 1.0)
            for kid, k in enumerate(s2.targets):
                save[f"rast_{rname}_k{kid}"] = s2.Vs[k]
                save[f"rast_{rname}_k{kid}_key"] = k
            save[f"rast_{rname}_Vglobal"] = s2.V
            save[f"rast_{rname}_room"] = json.dumps(rooms[rname] if rname in rooms else
                                                    json.load(open(f"rooms/{rname}.json")))
            save[f"dens_{rname}_xy"] = np.array([ag.position() for ag in s2.agents])
            s2.agents[0].status = False
            save[f"dens_{rname}_status"] = np.array([ag.status for ag in s2.agents], dtype=np.uint8)
            save[f"dens_{rname}"] = s2.gaussian_density(s2.sigma_convolution)
            save[f"init_{rname}_vdes"] = np.array([ag.v_des for ag in s2.agents])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print("wrote", name)


def run_golden(name, room, T, recompute=False, rooms=None, seed=0):
    """Whole `simulation.run()` of the reference: trajectories, evacuation times, exit order."""
    with rh.RefEnv(rooms) as env:
        np.random.seed(seed)
        with env.silence():
            simu = env.simulations.simulation(room, T, recompute)
            simu.run()
        N = simu.N
        steps = max(len(a.traj) for a in simu.agents)
        traj = np.full((N, steps, 4), np.nan)
        for i, a in enumerate(simu.agents):
            traj[i, :len(a.traj), :2] = np.array(a.traj)
            traj[i, :len(a.vels), 2:] = np.array(a.vels)
        save = {"versions": versions(), "T": T, "seed": seed, "recompute": int(recompute),
                "room": json.dumps(rooms[room] if rooms and room in rooms else json.load(open(f"rooms/{room}.json"))),
                "traj": traj, "agent_time": np.array([a.time for a in simu.agents]),
                "status": np.array([a.status for a in simu.agents], dtype=np.uint8), "time": simu.time,
                "simu_step": simu.simu_step, "inside": simu.inside, "n_history": len(simu.history),
                "v_des": np.array([a.v_des for a in simu.agents]),
                "nt_opt_final": np.array([simu.targets[k].nt_opt for k in simu.targets])}
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
        print("wrote", name, "steps", simu.simu_step, "time", simu.time, "inside", simu.inside)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    os.chdir(os.path.dirname(OUT.rstrip("/")) + "/..")
    rooms = {"small": SMALL_ROOM, "dense": DENSE_ROOM}
    unit_golden()
    hjb_golden("hjb_room_test_T3", "room_test", 3.0)
    hjb_golden("hjb_small_T2_density", "small", 2.0, rooms=rooms, with_density=True)
    hjb_golden("hjb_exit_opposite_T1", "exit_opposite", 1.0)
    gcfm_golden("gcfm_dense", "dense", 6.0, rooms=rooms, n_steps=160, field_every=8)
    gcfm_golden("gcfm_small", "small", 6.0, rooms=rooms, n_steps=250, field_every=10)
    run_golden("run_room_test_T30", "room_test", 30.0)
    run_golden("run_room_test_T4", "room_test", 4.0)
    run_golden("run_exit_opposite_T4_recompute", "exit_opposite", 4.0, recompute=True)
    # down-scaled BASELINE configs[2] (metro station: 4 boxes / 4 target sets, wall with holes, pillars) and configs[3]
    # (slalom cylinder field); the full-size rooms are generated by optimal_crowds_b200/synthetic.py
    hjb_golden("hjb_metro_station_T1", "metro_station", 1.0)
    run_golden("run_metro_station_T3", "metro_station", 3.0)
    run_golden("run_slalom_T4", "slalom", 4.0, seed=3)
