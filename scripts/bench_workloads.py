"""bench.py --workload metro | ensemble: BASELINE.json configs[2] and configs[4] measured through the drop-in API.

metro     4096 x 4096 grid, 10k agents, 4 boxes with 4 distinct target sets (= 4 HJB keys): ``simulation(room, T)`` on
          one GPU (keys solved one after the other, like the reference's loop simulations.py:424-425) or
          ``simulation(room, T, shard_keys=True)`` under torchrun (one key per GPU at N = 4, no data-path collective
          in the solve; the GCFM step merges the per-agent terms of the ranks bit-exactly).
ensemble  1024 independent members (512 x 512 grid, 1000 agents each, seeds 0..1023) sharded round-robin over the GPUs,
          each rank running its members in waves: batched HJB solves + interleaved GCFM steps (ensemble.py).
Both print ONE JSON line with the keys of bench.py's headline line.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_sample(room, nx, ny, T):
    """the C port of the reference's solve on a down-scaled room of the same layout (bounded CPU sample)"""
    from oracle import cpu_oracle as co
    L, H = room["room_length"], room["room_height"]
    X, Y = np.linspace(0, L, nx), np.linspace(0, H, ny)
    key = sorted({' or '.join(b[5:]) for b in room["initial_boxes"].values()})[0].split(' or ')
    V = co.create_potential(X, Y, list(room["walls"].values()), list(room["holes"].values()),
                            list(room["cylinders"].values()), [room["targets"][t] for t in key])
    V[V < 0] = -100; V[V > 0] = 1
    nt = round(T / 0.02)
    t0 = time.perf_counter()
    _, st, _, _ = co.hjb_solve(V, None, T, nt)
    dt = time.perf_counter() - t0
    return st["nfev"] * nx * ny / dt / 1e9, f"same layout at {nx}x{ny} nodes, one target set, T={T} (nfev={st['nfev']}), {dt:.1f} s"


def _common(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    import bench

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allred(vals, op="max"):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    t_start = time.perf_counter()

    def note(msg):
        if rank == 0:
            print(f"[bench {time.perf_counter() - t_start:7.1f}s] {msg}", file=sys.stderr, flush=True)

    return bench, barrier, allred, note


def run_metro(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    from optimal_crowds_b200 import _lib, simulations, synthetic
    bench, barrier, allred, note = _common(args, world, rank, local_rank)
    W, K = max(args.warmup, 0), args.steps
    n = 4096 if args.nx == bench.BAND_NX else args.nx
    agents = 10000 if args.agents == bench.BAND_AGENTS else args.agents
    T = args.T if args.T is not None else 2.0
    bench.pin_numa(local_rank)
    room = synthetic.metro_room(n, agents)
    np.random.seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        simu = simulations.simulation(room, T, record=world == 1, shard_keys=world > 1)
    mine = [k for k, o in simu.targets.items() if o.owned]
    note(f"metro: N={simu.N} agents, grid {simu.Ny}x{simu.Nx}, keys={list(simu.targets)}, this rank solves {mine}")
    for o in simu.targets.values():
        o._prm.profile = 1
    cells = simu.Ny * simu.Nx
    m_host = torch.from_numpy(bench.smooth_density(simu.Nx, simu.Ny, 0)).pin_memory()
    m_dev = m_host.to("cuda")

    def solve_all(m):
        nfev, cms, cby, cn = 0, 0.0, 0.0, 0
        with contextlib.redirect_stdout(io.StringIO()):
            for key in mine:                       # simulations.py:424-425: one solve per target set
                o = simu.targets[key]
                o.compute_optimal_velocity(0.0, m)
                st = o.last_stats
                nfev += st["nfev"]; cms += st["cls_ms"][0]; cby += st["cls_bytes"][0]; cn += st["cls_launches"][0]
        return nfev, cms, cby, cn

    for _ in range(W):
        solve_all(m_dev)
    barrier()
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = np.zeros(4)
    barrier()
    ev0.record()
    for _ in range(K):
        tot += solve_all(m_dev)
    ev1.record()
    barrier()
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_max, = allred([ev0.elapsed_time(ev1)])
    nfev_all, = allred([tot[0]], "sum")
    value = nfev_all * cells / (ms_max * 1e-3) / 1e9
    note(f"metro device-resident: {value:.1f} Gcu/s, {ms_max / K:.1f} ms per pass over the keys")
    # e2e: the density as a pinned host array through the reference-facing call, checksum read back
    def checksum():
        return sum(float(simu.targets[k].d_phi[simu.targets[k].nt_opt - 1].sum().item()) for k in mine)
    chk = 0.0
    solve_all(m_host.numpy()); checksum()
    barrier()
    t0 = time.perf_counter()
    nf = 0
    for _ in range(K):
        nf += solve_all(m_host.numpy())[0]
        chk = checksum()
    barrier()
    e2e_ms, = allred([(time.perf_counter() - t0) * 1e3])
    nf_all, = allred([nf], "sum")
    e2e_value = nf_all * cells / (e2e_ms * 1e-3) / 1e9
    # parity invariant at full size: the field of a key must not depend on who else shares the GPU -- re-solve with a
    # fixed reduction order and compare the t = 0 slice checksum of two solves bit for bit (determinism), and check the
    # door cells keep phi = 1's growth sign (phi > 0 everywhere, finite)
    inv = {"deterministic": True, "finite_positive": True}
    if mine:   # (a rank beyond the number of target sets owns none)
        o = simu.targets[mine[0]]
        a = o.d_phi[o.nt_opt - 1].clone()
        solve_all(m_host.numpy())
        inv = {"deterministic": bool(torch.equal(a, o.d_phi[o.nt_opt - 1])),
               "finite_positive": bool(torch.isfinite(a).all().item() and (a > 0).all().item())}
        del a
    else:
        solve_all(m_host.numpy())
    flags = allred([0.0 if inv["deterministic"] else 1.0, 0.0 if inv["finite_positive"] else 1.0])
    inv = {"deterministic": flags[0] == 0.0, "finite_positive": flags[1] == 0.0}
    # GCFM
    fp64_peak = simu._ctx.fp64_peak()
    with contextlib.redirect_stdout(io.StringIO()):
        simu._solve_all()
    agent_steps, dev_ms, pairs = 0, 0.0, 0
    if world == 1:   # through the block call of simulation.run() (bench.py does the same for the slalom workload)
        simu.advance(3)
        barrier()
        g0 = time.perf_counter()
        n0, s0 = int(simu._h_status.sum()), simu.simu_step
        done = simu.advance(args.gcfm_steps)
        ex = simu._exit_step[simu._exit_step >= s0]
        agent_steps = sum(n0 - int((ex < s0 + k).sum()) for k in range(done))
        dev_ms, pairs = simu.last_run_stats["device_ms"], simu.last_run_stats["pairs"]
        args.gcfm_steps = max(done, 1)
    else:
        for _ in range(3):
            simu.step(simu.dt)
        barrier()
        g0 = time.perf_counter()
        for _ in range(args.gcfm_steps):
            agent_steps += int(simu._h_status.sum())
            simu.step(simu.dt)
            dev_ms += simu._ctx.gcfm_last_ms(); pairs += simu._ctx.gcfm_last_pairs()
    barrier()
    g_wall, g_dev = allred([time.perf_counter() - g0, dev_ms * 1e-3])
    if rank == 0:
        peak, peak_src = bench.load_peaks()
        achieved = tot[2] / (tot[1] * 1e-3) / 1e9 if tot[1] > 0 else 0.0
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            g, sample = _cpu_sample(synthetic.metro_room(1024, 100), 1024, 1024, 0.5)
            cpu = {"value": g, "unit": "Gcell-updates/s", "cores": 1, "kind": "port", "sample": sample}
        line = {"metric": "hjb_gcell_updates_per_s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"metro_station (BASELINE configs[2]) {simu.Nx}x{simu.Ny} grid, {simu.N} agents, "
                                       f"{len(simu.targets)} target sets (HJB keys), wall with 8 holes, T={T} (nt={round(T / 0.02)})",
                           "parallelism": "1 GPU, keys solved in turn" if world == 1 else
                           f"target sets dealt round-robin to {world} GPUs (simulation(..., shard_keys=True)); no collective in "
                           "the solve; the GCFM step merges per-agent terms with one all-reduce(max) of bit patterns",
                           "l2": "inputs larger than L2 (each field 134 MB > 126 MB)", "field_storage": "phi (the API default)",
                           "api": "simulations.simulation(room, T).targets[key].compute_optimal_velocity(t, m) for every key",
                           "invariants": inv},
                "e2e": {"value": e2e_value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": int(m_host.numel() * 8) * len(simu.targets),
                        "d2h_bytes_per_step": 8 * len(simu.targets), "checksum": chk},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "fused::hjb_fused_kernel<NE>", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                             "launches": int(tot[3]), "avg_launch_ms": float(tot[1] / max(tot[3], 1)),
                             "share_of_step": float(tot[1] / (ev0.elapsed_time(ev1)))},
                "cpu_baseline": cpu,
                "gcfm": {"metric": "gcfm_agent_steps_per_s", "value": agent_steps / g_dev, "unit": "agent-steps/s",
                         "e2e_value": agent_steps / g_wall, "agents": simu.N, "pairs_per_step": pairs / args.gcfm_steps,
                         "roofline": {"bound": "fp64", "unit": "TFLOP/s", "peak": fp64_peak,
                                      "achieved": pairs * bench.FP64_INST_PER_PAIR * 2.0 / g_dev / 1e12,
                                      "frac": pairs * bench.FP64_INST_PER_PAIR * 2.0 / g_dev / 1e12 / fp64_peak}}}
        bench.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_ensemble(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    from optimal_crowds_b200 import _lib, ensemble, synthetic
    bench, barrier, allred, note = _common(args, world, rank, local_rank)
    W, K = max(args.warmup, 0), args.steps
    T = args.T if args.T is not None else 2.0
    n_rooms = args.rooms
    room = synthetic.ensemble_room(512, 1000)
    ctx = _lib.Context(room["room_length"], room["room_height"], 0.05)
    fp64_peak = ctx.fp64_peak()
    ctx.close()

    def one_pass():
        ens = ensemble.ensemble(room, T, list(range(n_rooms)), rank=rank, world=world, max_wave=128, chunk_rows="wave")
        t0 = time.perf_counter()
        res = ens.run(gather=False)
        torch.cuda.synchronize()
        return ens, res, (time.perf_counter() - t0) * 1e3

    for _ in range(W):
        one_pass()
    barrier()
    sampler = bench.ClockSampler(local_rank, interval_ms=1000)  # (an ensemble pass makes ~10^5 driver calls: sample sparsely)
    if rank == 0:
        sampler.start()
    _lib.launch_count(reset=True)
    acc = dict(hjb_ms=0.0, gcfm_ms=0.0, build_ms=0.0, agent_steps=0, cell_updates=0, wall_ms=0.0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        ens, res, wall = one_pass()
        for k in ("hjb_ms", "gcfm_ms", "build_ms", "agent_steps", "cell_updates"):
            acc[k] += ens.stats[k]
        acc["wall_ms"] += wall
        note("pass: " + ", ".join(f"{k} {ens.stats[k]:.0f}" for k in ("build_ms", "hjb_ms", "hjb_call_ms", "gcfm_ms", "launch_ms", "finish_ms"))
             + f", wall {wall:.0f} ms")
    barrier()
    total_ms = (time.perf_counter() - t0) * 1e3
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    hjb_ms, gcfm_ms, wall_ms = allred([acc["hjb_ms"], acc["gcfm_ms"], total_ms])
    cu, ast_ = allred([acc["cell_updates"], acc["agent_steps"]], "sum")
    evac = [r["evac_time"] for r in res.values()]
    if rank == 0:
        peak, peak_src = bench.load_peaks()
        value = cu / (hjb_ms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            g, sample = _cpu_sample(room, 512, 512, 1.0)
            cpu = {"value": g, "unit": "Gcell-updates/s", "cores": 1, "kind": "port", "sample": sample}
        line = {"metric": "hjb_gcell_updates_per_s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": wall_ms / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"ensemble (BASELINE configs[4]) of {n_rooms} independent members, 512x512 grid, 1000 "
                                       f"agents each (2.5 ped/m^2), seeds 0..{n_rooms - 1}, T={T}: batched HJB solve + the whole "
                                       "GCFM run of every member",
                           "parallelism": f"members round-robin over {world} GPU(s), waves of <= 128 members per GPU, no collective",
                           "api": "optimal_crowds_b200.ensemble.ensemble(room, T, seeds).run()",
                           "first_evac_times": evac[:4]},
                "e2e": {"value": cu / (wall_ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": int(n_rooms * 1000 * 8 * 5),
                        "note": "cell-updates over the wall clock of the whole pass (crowd placement, rasterisation, solve, "
                                "GCFM run, result read-back); rooms are rasterised on the device from their JSON shapes"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "fused::hjb_fused_kernel<NE> (batched over the members)",
                             "achieved": cu / world / 6.0 * 64.0 / (hjb_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": cu / world / 6.0 * 64.0 / (hjb_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                             "note": "per GPU; ~64 B per cell and attempt (40 B + ~2.6 phi slices x 8 B) over the whole batched-solve time "
                                     "(controller round trips included): small grids are latency-, not bandwidth-bound"},
                "cpu_baseline": cpu,
                "gcfm": {"metric": "gcfm_agent_steps_per_s", "value": ast_ / (gcfm_ms * 1e-3), "unit": "agent-steps/s",
                         "e2e_value": ast_ / (wall_ms * 1e-3), "members": n_rooms, "fp64_peak_tflops": fp64_peak}}
        bench.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0
