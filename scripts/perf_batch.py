"""Times oc_hjb_solve_batch (ensemble HJB) for R rooms of 512^2 on one context (development aid; run under gpurun).
usage: perf_batch.py [rooms] [T]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimal_crowds_b200 import _lib, synthetic

R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
cfg = json.load(open(os.path.join(os.path.dirname(_lib.__file__), "config.json")))
room = synthetic.ensemble_room(512, 1000)
ctx = _lib.Context(room["room_length"], room["room_height"], 0.05)
V = ctx.rasterise([], [], list(room["cylinders"].values()), list(room["targets"].values()), remap=True)
nt = round(T / 0.02)
prm = _lib.hjb_params(cfg, fused=1)
phis = [ctx.empty(nt, ctx.Ny, ctx.Nx) for _ in range(R)]
rng = np.random.RandomState(0)
ms_ = [ctx.to_device(rng.uniform(0, 1, (ctx.Ny, ctx.Nx))) for _ in range(R)]
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = ctx.hjb_solve_batch([V] * R, ms_, prm, T, nt, out_phi=phis)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nfev = sum(r["stats"]["nfev"] for r in res)
    print(f"{os.path.basename(_lib.LIB_PATH)} rooms {R} T {T}: {dt*1e3:.1f} ms, nfev/room {nfev/R:.0f}, "
          f"{nfev*ctx.Ny*ctx.Nx/dt/1e9:.1f} Gcu/s, {dt*1e6/((nfev-2*R)/6):.1f} us per attempt", flush=True)
