"""configs[4] figure: an ensemble of 512^2 rooms with 1k agents each on one GPU (run under gpurun).
usage: perf_ensemble.py [members] [T] [max_wave]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimal_crowds_b200 import ensemble, synthetic

members = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
max_wave = int(sys.argv[3]) if len(sys.argv) > 3 else 64
room = synthetic.ensemble_room(512, 1000)
for chunk, ctas in ((64, 0), (64, 16)):
    ens = ensemble.ensemble(room, T, list(range(members)), chunk_rows=chunk, max_wave=max_wave)
    ens.sweep_ctas = ctas
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = ens.run()
    dt = time.perf_counter() - t0
    st = ens.stats
    done = sum(1 for r in res.values() if r["inside"] == 0)
    print(f"chunk_rows={chunk} sweep_ctas={ctas}: {members} rooms (512^2, 1000 agents, T={T}) in {dt:.2f} s = {members/dt:.2f} rooms/s; waves={st['waves']} "
          f"build {st['build_ms']/1e3:.2f} s | HJB {st['hjb_ms']/1e3:.2f} s = {st['cell_updates']/st['hjb_ms']/1e6:.1f} Gcu/s | "
          f"GCFM {st['gcfm_ms']/1e3:.2f} s = {st['agent_steps']/st['gcfm_ms']/1e3:.2f} M agent-steps/s; "
          f"(host launch {st['launch_ms']/1e3:.2f} s, finish {st['finish_ms']/1e3:.2f} s) "
          f"evacuated {done}/{members}, mean steps {np.mean([r['steps'] for r in res.values()]):.0f}", flush=True)
