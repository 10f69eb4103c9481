"""development aid: agent-steps/s through simulation.advance (the block call of run()) vs simulation.step"""
import contextlib, io, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_crowds_b200 import simulations, synthetic
for agents in [int(a) for a in sys.argv[1:]] or [12500]:
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        simu = simulations.simulation(synthetic.slalom_room(16384, 2048, agents=agents), 2.0)
        simu._solve_all()
    simu.advance(5)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = simu.advance(40)
    torch.cuda.synchronize(); w = time.perf_counter() - t0
    st = simu.last_run_stats
    print(f"N={simu.N}: advance({n}) {w/n*1e3:.3f} ms/step -> {simu.N*n/w/1e6:.2f} M agent-steps/s e2e; device {st['device_ms']/n:.3f} ms/step "
          f"-> {simu.N*n/st['device_ms']/1e3:.2f} M/s; ratio {st['device_ms']/n/(w/n*1e3):.2f}", flush=True)
    del simu
    torch.cuda.empty_cache()
