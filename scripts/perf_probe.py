"""Quick HJB perf probe (development aid): synthetic cylinder-lattice room built with torch ops."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_crowds_b200 import _lib

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from perf_probe_common import make_V

cfg = json.load(open(os.path.join(os.path.dirname(_lib.__file__), "config.json")))
for (Ny, Nx, T) in [(4096, 4096, 1.0), (2048, 16384, 1.0)]:
    L, H = (Nx - 1) * 0.05 + 0.025, (Ny - 1) * 0.05 + 0.025
    ctx = _lib.Context(L, H, 0.05)
    V = make_V(ctx)
    prm = _lib.hjb_params(cfg)
    nt = round(T / 0.02)
    vx = ctx.empty(nt - 1, Ny - 2, Nx - 2); vy = ctx.empty(nt - 1, Ny - 2, Nx - 2)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        res = ctx.hjb_solve(V, None, prm, T, nt, out_vx=vx, out_vy=vy)
        torch.cuda.synchronize(); dt = time.time() - t0
        st = res["stats"]
        cu = st["nfev"] * Ny * Nx
        print(f"{Ny}x{Nx} T={T} nfev={st['nfev']} acc={st['n_accepted']} rej={st['n_rejected']} launches={st['launches']} "
              f"gpu_ms={st['gpu_ms']:.2f} wall_ms={dt*1e3:.2f} Gcu/s={cu/st['gpu_ms']/1e6:.2f} "
              f"eqGB/s(52B)={cu*52/st['gpu_ms']/1e6:.0f}", flush=True)
    del vx, vy, V; ctx.close(); torch.cuda.empty_cache()
