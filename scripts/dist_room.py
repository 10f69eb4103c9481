"""BASELINE configs[3] as ONE room on N GPUs (run under torchrun): the 16384 x (2048 N) slalom room with 12.5k N agents,
row-decomposed: band-wise HJB solve (NCCL halo exchange), then GCFM steps through simulation.step with the field
samples / wall forces evaluated by the rank that owns the agent's rows and a replicated sweep.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 scripts/dist_room.py
"""
import contextlib, io, json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_crowds_b200 import simulations, synthetic

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nx = int(os.environ.get("OC_NX", "16384")); ny = int(os.environ.get("OC_BAND_NY", "2048")) * world
agents = int(os.environ.get("OC_AGENTS", "12500")) * world
T, steps = float(os.environ.get("OC_T", "2.0")), int(os.environ.get("OC_STEPS", "20"))
t0 = time.perf_counter()
room = synthetic.slalom_room(nx, ny, agents=agents)
np.random.seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    simu = simulations.simulation(room, T, recompute=False, record=False, field_storage="phi", band=True)
torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    simu._solve_all()
torch.cuda.synchronize(); dist.barrier(); t2 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    simu._solve_all()
torch.cuda.synchronize(); dist.barrier(); t3 = time.perf_counter()
for _ in range(3):
    simu.step(simu.dt)
torch.cuda.synchronize(); dist.barrier(); t4 = time.perf_counter()
n_steps, dev_ms = 0, 0.0
for _ in range(steps):
    n_steps += int(simu._h_status.sum())
    simu.step(simu.dt)
    dev_ms += simu._ctx.gcfm_last_ms()
torch.cuda.synchronize(); dist.barrier(); t5 = time.perf_counter()
chk = torch.tensor([float(np.sum(simu._state["x"].cpu().numpy()))], dtype=torch.float64, device="cuda")
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
st = list(simu.targets.values())[0].last_stats
if rank == 0:
    cells = nx * ny
    print(json.dumps({"what": "configs[3] as one row-decomposed room", "n_gpus": world, "grid": [ny, nx], "agents": simu.N, "T": T,
                      "init_s": round(t1 - t0, 2), "hjb_first_solve_s": round(t2 - t1, 3), "hjb_solve_s": round(t3 - t2, 4),
                      "hjb_gcell_updates_per_s": round(st["nfev"] * cells / (t3 - t2) / 1e9, 1), "nfev": st["nfev"],
                      "gcfm_steps": steps, "gcfm_ms_per_step_e2e": round((t5 - t4) / steps * 1e3, 3),
                      "gcfm_ms_per_step_device": round(dev_ms / steps, 3),
                      "gcfm_agent_steps_per_s_e2e": round(n_steps / (t5 - t4)), "gcfm_agent_steps_per_s_device": round(n_steps / dev_ms * 1e3),
                      "replicas_agree_bitwise": bool(lo.item() == hi.item()), "inside": int(simu.inside)}), flush=True)
dist.destroy_process_group()
