"""Multi-GPU parity check (run under torchrun, one rank per GPU):
the NCCL row-band solve must equal, bit for bit, the same rank's rows of a virtual-band solve of the whole grid
done locally on that GPU (which the single-GPU tests tie to the oracle).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_crowds_b200 import _lib, dist as ocd

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = json.load(open(os.path.join(os.path.dirname(_lib.__file__), "config.json")))
Ny, Nx, T = 128 * world, 520, 0.6
L, H = (Nx - 1) * 0.05 + 0.025, (Ny - 1) * 0.05 + 0.025
rng = np.random.RandomState(11)
V = np.zeros((Ny, Nx)); V[0] = V[-1] = -100; V[:, 0] = V[:, -1] = -100
V[120:136, 40:300] = -100; V[Ny // 2 - 10: Ny // 2 + 10, -1] = 1.0; V[0, 20:40] = 1.0
m = rng.uniform(0, 1, (Ny, Nx))
nt = round(T / 0.02)
ctx = _lib.Context(L, H, 0.05)
own0, own1 = ocd.init_context(ctx)
prm = _lib.hjb_params(cfg, fused=1, chunk_rows=32)
res = ctx.hjb_solve_band(ctx.to_device(V[own0:own1]), ctx.to_device(m[own0:own1]), prm, T, nt, own=(own0, own1),
                         want_vel=True, trace=True)
ctx2 = _lib.Context(L, H, 0.05)
ref = ctx2.hjb_solve_band(ctx2.to_device(V), ctx2.to_device(m), prm, T, nt, n_virtual=world, want_vel=True, trace=True)
ok = np.array_equal(res["trace_h"], ref["trace_h"]) and np.array_equal(res["trace_err"], ref["trace_err"])
phi_b = res["phi"][:, 1:-1]                       # owned rows of the band array (halo rows stripped)
ok = ok and torch.equal(phi_b, ref["phi"][:, own0:own1])
lo, hi = max(own0, 1), min(own1, Ny - 1)          # interior rows owned by this rank
ok = ok and torch.equal(res["vx"][:, lo - own0: hi - own0], ref["vx"][:, lo - 1: hi - 1])
ok = ok and torch.equal(res["vy"][:, lo - own0: hi - own0], ref["vy"][:, lo - 1: hi - 1])
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "OK" if flag.item() == 1 else "MISMATCH", "world", world, "nfev", res["stats"]["nfev"],
          "attempts", len(res["trace_h"]), flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
