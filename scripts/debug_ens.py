import os, sys, time, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimal_crowds_b200 import ensemble, synthetic, simulations

mode = sys.argv[1]
members, T = 8, 1.0
room = synthetic.ensemble_room(512, 1000)
ens = ensemble.ensemble(room, T, list(range(members)))
sims = [ens._build(i) for i in range(members)]
torch.cuda.synchronize()
if mode in ("batch", "both"):
    ens._solve_wave(sims)
    for i, s in enumerate(sims):
        o = list(s.targets.values())[0]
        phi_b = o.d_phi.clone()
        with contextlib.redirect_stdout(io.StringIO()):
            o.compute_optimal_velocity(0.0, None)
        print("member", i, "batch==single:", torch.equal(phi_b, o.d_phi), "finite:", bool(torch.isfinite(phi_b).all()), flush=True)
else:
    for s in sims:
        with contextlib.redirect_stdout(io.StringIO()):
            s._solve_all()
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in sims]
for step in range(40):
    if mode in ("conc", "both"):
        l = []
        for q, s in enumerate(sims):
            with torch.cuda.stream(streams[q]):
                l.append(s._step_launch(s.dt))
        for s, x in zip(sims, l):
            s._step_finish(x)
    else:
        for s in sims:
            s.step(s.dt)
torch.cuda.synchronize()
print(mode, "ok", [s.inside for s in sims])
