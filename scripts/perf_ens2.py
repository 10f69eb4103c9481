"""development aid: where does an ensemble pass spend its time (run under gpurun)"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimal_crowds_b200 import ensemble, synthetic
R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
room = synthetic.ensemble_room(512, 1000)
for it in range(3):
    ens = ensemble.ensemble(room, T, list(range(R)), max_wave=128)
    pr = cProfile.Profile() if it == 2 else None
    t0 = time.perf_counter()
    if pr: pr.enable()
    res = ens.run(gather=False)
    if pr: pr.disable()
    dt = time.perf_counter() - t0
    st = ens.stats
    print(f"pass {it}: {dt:.2f} s build {st['build_ms']:.0f} ms hjb {st['hjb_ms']:.0f} ms ({st['cell_updates']/st['hjb_ms']/1e6:.1f} Gcu/s) "
          f"gcfm {st['gcfm_ms']:.0f} ms ({st['agent_steps']/st['gcfm_ms']/1e3:.2f} M agent-steps/s) launch {st['launch_ms']:.0f} finish {st['finish_ms']:.0f}", flush=True)
    if pr:
        pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
