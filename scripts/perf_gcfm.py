"""GCFM throughput probe: agent-steps/s of oc_gcfm_step vs crowd size on the slalom band (development aid)."""
import contextlib, io, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_crowds_b200 import simulations, synthetic

cases = ((12500, 2048), (50000, 2048), (100000, 2048))
if len(sys.argv) > 1:
    cases = tuple((int(a), 2048) for a in sys.argv[1:])
for agents, ny in cases:
    np.random.seed(0)
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        simu = simulations.simulation(synthetic.slalom_room(16384, ny, agents=agents), 1.0, record=False,
                                      field_storage="phi", fused=1)
        simu._solve_all()
    for kv in filter(None, os.environ.get("OC_KNOBS", "").split(",")):  # e.g. OC_KNOBS=gcfm_margin_mm=1000,gcfm_fov_cull=0
        k, v = kv.split("=")
        simu._ctx.set_int(k, int(v))
    t1 = time.time()
    for _ in range(3):
        simu.step(simu.dt)
    ms, n = 0.0, 0
    w0 = time.time()
    for _ in range(10):
        n += int(simu._h_status.sum())
        simu.step(simu.dt)
        ms += simu._ctx.gcfm_last_ms()
    w = time.time() - w0
    print(os.environ.get("OC_KNOBS", ""), f"redos {simu._ctx.gcfm_last_redos() if hasattr(simu._ctx, 'gcfm_last_redos') else '?'}", end=" ")
    print(f"N={simu.N} init {t1-t0:.1f}s  device {ms/10:.3f} ms/step -> {n/ms/1e3:.2f} M agent-steps/s ; "
          f"through simulation.step {w/10*1e3:.3f} ms/step -> {n/w/1e6:.2f} M/s", flush=True)
    del simu
    torch.cuda.empty_cache()
