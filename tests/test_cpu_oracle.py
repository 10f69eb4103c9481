"""CPU tests: the oracle (oracle/*.c) against the golden vectors produced by the UNMODIFIED reference
(oracle/make_goldens.py).  This is what pins the oracle; the GPU tests then compare CUDA with the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import golden, room_grid
from oracle import cpu_oracle as co


def _ulp(a, b):
    return np.max(np.abs(a - b) / np.spacing(np.abs(b)))


def test_elementary_functions_within_2ulp_of_numpy():
    rng = np.random.RandomState(0)
    x = rng.uniform(-100, 100, 200000)
    assert _ulp(co.math_fn("exp", x), np.exp(x)) <= 1
    x = rng.uniform(-7, 7, 200000)
    assert _ulp(co.math_fn("sin", x), np.sin(x)) <= 1 and _ulp(co.math_fn("cos", x), np.cos(x)) <= 1
    y, x = rng.normal(size=200000), rng.normal(size=200000)
    assert _ulp(co.math_fn("atan2", y, x), np.arctan2(y, x)) <= 2
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf])
    Y, X = [a.ravel().copy() for a in np.meshgrid(sp, sp)]
    a, b = co.math_fn("atan2", Y, X), np.arctan2(Y, X)
    assert np.array_equal(a, b) and np.array_equal(np.signbit(a), np.signbit(b))
    assert co.math_fn("exp", np.array([800.0]))[0] == np.inf and co.math_fn("exp", np.array([-800.0]))[0] == 0.0


@pytest.mark.parametrize("name", ["hjb_room_test_T3", "hjb_small_T2_density", "hjb_exit_opposite_T1",
                                  "hjb_metro_station_T1"])
def test_hjb_oracle_vs_reference(name):
    g = golden(name)
    T = float(g["T"])
    for kid in range(int(g["n_keys"])):
        V = g[f"k{kid}_V"]; Ny, Nx = V.shape
        m = g[f"k{kid}_m"] if f"k{kid}_m" in g else None
        nt = int(g[f"k{kid}_nt"])
        phi, st, th, te = co.hjb_solve(V, m, T, nt)
        assert st["nfev"] == int(g[f"k{kid}_nfev"]) and st["status"] == 0 and st["n_out"] == nt
        np.testing.assert_allclose(st["h0"], float(g[f"k{kid}_h0"]), rtol=1e-13)
        np.testing.assert_allclose(th, g[f"k{kid}_attempt_h"], rtol=1e-12)
        np.testing.assert_allclose(te, g[f"k{kid}_attempt_err"], rtol=1e-10)
        vx, vy = co.fill_field(phi, Ny, Nx)
        sl = g[f"k{kid}_slices"]
        for q, s in enumerate(sl):
            np.testing.assert_allclose(phi[nt - 1 - s], g[f"k{kid}_phi"][q], rtol=1e-12)
        assert np.abs(vx[sl] - g[f"k{kid}_vx"]).max() < 1e-11
        assert np.abs(vy[sl] - g[f"k{kid}_vy"]).max() < 1e-11
        np.testing.assert_allclose(vx.sum(axis=(1, 2)), g[f"k{kid}_vx_sum"], atol=1e-8)
        np.testing.assert_allclose(vy.sum(axis=(1, 2)), g[f"k{kid}_vy_sum"], atol=1e-8)


def test_hjb_rhs_bit_exact_vs_numpy_expression():
    """the RHS is elementwise IEEE arithmetic in the reference's order -> bit-identical to numpy (optimals.py:144-164)."""
    rng = np.random.RandomState(1)
    Ny, Nx = 41, 57
    V = np.zeros((Ny, Nx)); V[0] = V[-1] = -100; V[:, 0] = V[:, -1] = -100; V[10:14, 20:30] = -100; V[20:24, -1] = 1
    m = rng.uniform(0, 1, (Ny, Nx)); phi = np.exp(rng.normal(size=(Ny, Nx)))
    sigma, mu, g, dx = 0.2, 5, -0.005, 0.05
    pt = np.empty((Ny + 2, Nx + 2)); pt[1:-1, 1:-1] = phi
    pt[0, :] = pt[2, :]; pt[-1, :] = pt[-3, :]; pt[:, -1] = pt[:, -3]; pt[:, 0] = pt[:, 2]
    lap = (pt[:-2, 1:-1] + pt[2:, 1:-1] + pt[1:-1, :-2] + pt[1:-1, 2:] - 4 * pt[1:-1, 1:-1]) / (dx * dx)
    ref = -0.5 * sigma ** 2 * lap - ((V + g * m) * pt[1:-1, 1:-1]) / (mu * sigma ** 2)
    ref[V < 0] = 0
    out = co.hjb_rhs(phi.ravel(), V, m, dx, dx, sigma, mu, g).reshape(Ny, Nx)
    assert np.array_equal(out, ref)


def _params(cfg, room):
    L, H, Ny, Nx, X, Y = room_grid(room)
    return co.gcfm_params(cfg, L, H, Ny, Nx), X, Y, Ny, Nx


def test_pair_wall_sampler_units_vs_reference(cfg):
    u = golden("units")
    room = json.loads(str(u["room"]))
    P, X, Y, Ny, Nx = _params(cfg, room)
    out = np.array([co.pair_force(P, u["pair_pi"][q], u["pair_vi"][q], u["pair_vdes"][q], u["pair_pj"][q],
                                  u["pair_vj"][q]) for q in range(len(u["pair_pi"]))])
    assert np.abs(out - u["pair_out"]).max() < 1e-12
    res = np.array([co.wall_force(P, X, Y, u["wall_V"], u["wall_p"][q], u["wall_v"][q], u["wall_vdes"][q])
                    for q in range(len(u["wall_p"]))])
    assert np.array_equal(res[:, 2].astype(np.int64), u["wall_ind"])  # np.argmin incl. tie-breaks
    assert np.abs(res[:, :2] - u["wall_out"]).max() < 1e-12
    k = co.KeyData(u["wall_V"], u["samp_vx"], u["samp_vy"], int(u["samp_nt_opt"]), [[0, 0, 1, 1]])
    n_err = 0
    for q in range(len(u["samp_p"])):
        ox, oy, bad = co.choose_velocity(P, k, u["samp_p"][q, 0], u["samp_p"][q, 1], int(u["samp_t"][q]))
        if u["samp_ok"][q]:
            assert bad == 0 and ox == u["samp_out"][q, 0] and oy == u["samp_out"][q, 1]
        else:
            assert bad == 1  # the reference raised IndexError here
            n_err += 1
    assert n_err > 0


@pytest.mark.parametrize("rname", ["room_test", "exit_opposite", "dense", "small"])
def test_rasteriser_and_density_vs_reference(rname):
    u = golden("units")
    room = json.loads(str(u[f"rast_{rname}_room"]))
    L, H, Ny, Nx, X, Y = room_grid(room)
    kid = 0
    while f"rast_{rname}_k{kid}" in u:
        key = str(u[f"rast_{rname}_k{kid}_key"])
        V = co.create_potential(X, Y, list(room["walls"].values()), list(room["holes"].values()),
                                list(room["cylinders"].values()), [room["targets"][t] for t in key.split(" or ")])
        V[V < 0] = -100; V[V > 0] = 1
        assert np.array_equal(V, u[f"rast_{rname}_k{kid}"])
        kid += 1
    xy = u[f"dens_{rname}_xy"]
    d = co.density(X, Y, u[f"rast_{rname}_Vglobal"], xy[:, 0], xy[:, 1], u[f"dens_{rname}_status"], 0.5)
    assert np.abs(d - u[f"dens_{rname}"]).max() < 1e-14


def _keys_for_step(g, room, Ny, Nx, s):
    keys = []
    for kid in range(int(g["n_keys"])):
        key = str(g[f"k{kid}_key"])
        doors = [room["targets"][t] for t in key.split(" or ")]
        vx = np.zeros((s + 1, Ny - 2, Nx - 2)); vy = np.zeros_like(vx)
        vx[s] = g[f"k{kid}_vx_step{s}"]; vy[s] = g[f"k{kid}_vy_step{s}"]
        keys.append(co.KeyData(g[f"k{kid}_V"], vx, vy, int(g[f"k{kid}_nt"]), doors))
    return keys


@pytest.mark.parametrize("name", ["gcfm_dense", "gcfm_small"])
def test_gcfm_teacher_forced_vs_reference(name, cfg):
    """O2 vs O1, one step from the reference's recorded state (two-oracle protocol)."""
    g = golden(name)
    room = json.loads(str(g["room"]))
    P, X, Y, Ny, Nx = _params(cfg, room)
    n_exits = 0
    for s in [int(q) for q in g["field_steps"]]:
        b, a = g["before"][s], g["after"][s]
        st = {k: np.ascontiguousarray(b[:, i]) for i, k in enumerate(("x", "y", "vx", "vy"))}
        st["time"] = np.ascontiguousarray(b[:, 5]); st["status"] = b[:, 4].astype(np.uint8)
        noise = g["noise"][s]; noise = noise[~np.isnan(noise[:, 0])]
        ex, bad, _ = co.gcfm_step(P, st, g["v_des"], g["agent_key"], _keys_for_step(g, room, Ny, Nx, s), X, Y,
                                  g["perm"][s], noise, s)
        out = np.column_stack([st[k] for k in ("x", "y", "vx", "vy")] + [st["status"].astype(float), st["time"]])
        assert bad == 0
        assert np.abs(out - a).max() < 1e-12
        assert np.array_equal(out[:, 4], a[:, 4])  # exits bit-exact
        n_exits += len(ex)
    assert n_exits == int((g["before"][0][:, 4].sum() - g["inside"][-1])) or name == "gcfm_dense"


@pytest.mark.parametrize("name,room_name,seed,horizon", [("run_metro_station_T3", "metro_station", 0, 60),
                                                         ("run_slalom_T4", "slalom", 3, 60)])
def test_whole_run_restatement_vs_reference_run(name, room_name, seed, horizon, cfg):
    """O2 against O1 over a free run (down-scaled BASELINE configs[2] and [3]: 4 boxes with 4 target sets, two of them
    multi-target, a wall pierced by holes, pillars / a cylinder field): crowd placement with the reference's RNG
    stream, one oracle HJB solve per target set, sequential sweeps -- trajectories within 1e-8 of the reference's
    over the first 60 steps (chaos separates any two implementations later)."""
    from optimal_crowds_b200 import _crowd
    from conftest import REPO, room_grid
    g = golden(name)
    room = json.loads(str(g["room"]))
    assert room == json.load(open(os.path.join(REPO, "rooms", room_name + ".json")))
    T = float(g["T"])
    L, H, Ny, Nx, X, Y = room_grid(room)
    P = co.gcfm_params(cfg, L, H, Ny, Nx)
    np.random.seed(seed)
    place = np.zeros((Ny, Nx))
    keys, kdata, xs, ys, vd, kid = [], [], [], [], [], []
    nt = round(T / cfg["dt"])
    for box in room["initial_boxes"].values():
        tg = box[5:]
        key = " or ".join(tg)
        if key not in keys:
            doors = [room["targets"][t] for t in tg]
            V = co.create_potential(X, Y, list(room["walls"].values()), list(room["holes"].values()),
                                    list(room["cylinders"].values()), doors)
            V[V < 0] = -100; V[V > 0] = 1
            phi, st, _, _ = co.hjb_solve(V, None, T, nt)
            assert st["status"] == 0
            vx, vy = co.fill_field(phi, Ny, Nx)
            keys.append(key)
            kdata.append(co.KeyData(V, vx, vy, nt, doors))
        a, b, c = _crowd.place_box(box, X, Y, place)
        xs.append(a); ys.append(b); vd.append(c); kid.append(np.full(len(a), keys.index(key), dtype=np.int32))
    xs, ys, vd, kid = map(np.concatenate, (xs, ys, vd, kid))
    traj = g["traj"]
    N = len(xs)
    assert N == traj.shape[0] and np.array_equal(vd, g["v_des"])
    assert np.array_equal(np.column_stack([xs, ys]), traj[:, 0, :2])          # identical initial crowd
    st = dict(x=xs.copy(), y=ys.copy(), vx=np.zeros(N), vy=np.zeros(N), time=np.zeros(N), status=np.ones(N, dtype=np.uint8))
    for s in range(horizon):
        perm = np.random.choice(np.arange(N), N, replace=False)
        noise = np.random.normal(size=(int(st["status"].sum()), 2))
        ex, bad, _ = co.gcfm_step(P, st, vd, kid, kdata, X, Y, perm, noise, s)
        assert bad == 0
        act = st["status"] == 1
        ref = traj[:, s + 1]
        assert np.abs(np.column_stack([st["x"], st["y"], st["vx"], st["vy"]])[act] - ref[act]).max() < 1e-8, f"step {s}"
