"""Multi-GPU parity check of a whole row-decomposed simulation (run under torchrun, one rank per GPU):
``simulation(room, T, band=True)`` -- band-wise HJB solve over NCCL, field samples / wall forces evaluated by the
rank that owns the agent's rows, bit-exact merge, replicated sweep -- must equal the single-GPU simulation of the
same room and seed bit for bit (positions, velocities, clocks, exit order) on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_gcfm_check.py
"""
import contextlib, io, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_crowds_b200 import simulations, synthetic

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nx, ny, agents, T, steps = 1024, 256 * world, 600, 1.0, int(os.environ.get("OC_STEPS", "40"))
room = synthetic.slalom_room(nx, ny, agents=agents, pitch=6.0, door_pitch=24.0)
recompute = os.environ.get("OC_RECOMPUTE", "0") == "1"


def run(band):
    np.random.seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        simu = simulations.simulation(room, T, recompute=recompute, record=False, field_storage="phi", chunk_rows=32, band=band)
        simu.recompute_step = 15
        simu._solve_all()
        for s in range(steps):
            if recompute and s > 0 and s % simu.recompute_step == 0:
                simu._solve_all()
            simu.step(simu.dt)
    simu._sync_host()
    return simu


a = run(True)
b = run(False)
if os.environ.get("OC_DIAG", "1") == "1":
    oa, ob = list(a.targets.values())[0], list(b.targets.values())[0]
    own0, own1 = a._band
    d = (oa.d_phi[:, 1:own1 - own0 + 1] - ob.d_phi[:, own0:own1]).abs().max().item()
    dh = (oa.d_phi[:, own1 - own0 + 1:own1 - own0 + 3] - ob.d_phi[:, own1:own1 + 2]).abs().max().item() if own1 + 2 <= oa.Ny else -1
    dl = (oa.d_phi[:, 0] - ob.d_phi[:, own0 - 1]).abs().max().item() if own0 > 0 else -1
    print(f"[rank {rank}] phi band vs single: owned rows max|diff| {d:.3e}, high halo {dh:.3e}, low halo {dl:.3e}; "
          f"state max|diff| {np.abs(a._h_now - b._h_now).max():.3e}; h0 {oa.last_stats['h0']!r} vs {ob.last_stats['h0']!r}", flush=True)
ok = np.array_equal(a._h_now, b._h_now) and np.array_equal(a._h_timev, b._h_timev) and a._exit_order == b._exit_order \
    and a.inside == b.inside
nfev_a = [o.last_stats["nfev"] for o in a.targets.values()]
nfev_b = [o.last_stats["nfev"] for o in b.targets.values()]
ok = ok and nfev_a == nfev_b
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    moved = float(np.abs(a._h_now[:, :2] - a._track[0][:, :2]).max())
    print("DIST_GCFM_CHECK", "OK" if flag.item() == 1 else "MISMATCH", "world", world, "agents", a.N, "steps", steps,
          "exits", len(a._exit_order), "max displacement", round(moved, 3), "nfev", nfev_a, "recompute", recompute, flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
