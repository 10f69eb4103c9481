"""Where does the end-to-end time of optimals.compute_optimal_velocity(t, m_host) go? (run under gpurun)"""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimal_crowds_b200 import _lib, simulations, synthetic

nx, ny, T = 16384, 2048, 2.0
room = synthetic.slalom_room(nx, ny, agents=1000)
with contextlib.redirect_stdout(io.StringIO()):
    simu = simulations.simulation(room, T, recompute=False, record=False, field_storage="phi", fused=1)
opt = simu.targets[list(simu.targets)[0]]
m_host = torch.zeros((ny, nx), dtype=torch.float64).pin_memory()
m_dev = m_host.to("cuda")
m_np = m_host.numpy()


def tm(label, fn, n=3):
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        print(f"{label:40s} #{i}: {(time.perf_counter() - t0) * 1e3:8.2f} ms   gpu_ms(solve)={opt.last_stats['gpu_ms'] if opt.last_stats else 0:.2f}", flush=True)


def quiet(fn):
    def g():
        with contextlib.redirect_stdout(io.StringIO()):
            fn()
    return g


tm("solve(m_dev)", quiet(lambda: opt.compute_optimal_velocity(0.0, m_dev)))
tm("solve(m_host numpy, pinned)", quiet(lambda: opt.compute_optimal_velocity(0.0, m_np)))
tm("solve(None)", quiet(lambda: opt.compute_optimal_velocity(0.0, None)))
tm("solve(m_dev) again", quiet(lambda: opt.compute_optimal_velocity(0.0, m_dev)))
tm("upload only", lambda: opt._ctx.upload(m_np, opt._d_m))
tm("checksum", lambda: float(opt.d_phi[opt.nt_opt - 1].sum().item()))
