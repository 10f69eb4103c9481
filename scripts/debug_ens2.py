import os, sys, time, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimal_crowds_b200 import ensemble, synthetic, simulations
room = synthetic.ensemble_room(512, 1000)
M = 32
ens = ensemble.ensemble(room, 1.0, list(range(M)), chunk_rows=64)
sims = [ens._build(i) for i in range(M)]
ens._solve_wave(sims)
streams = [torch.cuda.Stream() for _ in sims]
def rounds(n_mem, ctas, poll, steps=10):
    for s in sims[:n_mem]:
        s._ctx.set_int("gcfm_sweep_ctas", ctas); s._ctx.set_int("gcfm_poll_ns", poll)
    torch.cuda.synchronize(); t0 = time.perf_counter(); dev = 0.0
    for _ in range(steps):
        l = []
        for q, s in enumerate(sims[:n_mem]):
            with torch.cuda.stream(streams[q]):
                l.append(s._step_launch(s.dt))
        for s, x in zip(sims, l):
            s._step_finish(x); dev += s._ctx.gcfm_last_ms()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3, dev / steps / n_mem
for n_mem in (1, 2, 4, 8, 16, 32):
    for ctas, poll in ((0, 20), (16, 20), (16, 500), (0, 500), (4, 500), (16, 2000)):
        ms, dev = rounds(n_mem, ctas, poll)
        print(f"members={n_mem:3d} ctas={ctas:3d} poll={poll:5d} ns: {ms:8.2f} ms/round  ({ms/n_mem:6.2f} per member; device event time per member-step {dev:6.2f} ms)", flush=True)
