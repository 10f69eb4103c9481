"""development aid: host-side profile of simulation.step at 12.5 k agents (run under gpurun)"""
import cProfile, contextlib, io, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_crowds_b200 import simulations, synthetic
np.random.seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    simu = simulations.simulation(synthetic.slalom_room(16384, 2048, agents=int(sys.argv[1]) if len(sys.argv) > 1 else 12500), 1.0, record=False)
    simu._solve_all()
for _ in range(5):
    simu.step(simu.dt)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for _ in range(40):
    simu.step(simu.dt)
dt = time.perf_counter() - t0
pr.disable()
print(f"{dt/40*1e3:.3f} ms/step, lookahead hits {simu._rng.hits} misses {simu._rng.misses}")
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
