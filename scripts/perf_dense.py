"""development aid: GCFM step time of ONE dense room (BASELINE configs[4] member: 512 x 512 nodes, 1000 agents at 2.5 ped/m^2)"""
import contextlib, io, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_crowds_b200 import simulations, synthetic
np.random.seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    simu = simulations.simulation(synthetic.ensemble_room(512, int(sys.argv[1]) if len(sys.argv) > 1 else 1000), 2.0, record=False)
    simu._solve_all()
for _ in range(5):
    simu.step(simu.dt)
ms, n, pairs = 0.0, 0, 0
w0 = time.perf_counter()
for _ in range(30):
    n += int(simu._h_status.sum())
    simu.step(simu.dt)
    ms += simu._ctx.gcfm_last_ms()
    pairs += simu._ctx.gcfm_last_pairs()
w = time.perf_counter() - w0
print(os.environ.get("OC_KNOBS", ""), f"N={simu.N}: device {ms/30:.3f} ms/step -> {n/ms/1e3:.2f} M agent-steps/s; through step {w/30*1e3:.3f} ms; "
      f"pairs/step {pairs/30:.0f}; redos {simu._ctx.gcfm_last_redos()}")
