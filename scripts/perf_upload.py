"""H2D paths of the density input (run under gpurun): torch copy vs oc_upload, page-locked vs pageable source."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimal_crowds_b200 import _lib

ny, nx = 2048, 16384
ctx = _lib.Context((nx - 1) * 0.05 + 0.025, (ny - 1) * 0.05 + 0.025, 0.05)
pin = torch.zeros((ny, nx), dtype=torch.float64).pin_memory()
pag = np.zeros((ny, nx))
out = ctx.empty(ny, nx)


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


gb = pin.numel() * 8 / 1e9
for name, fn in (("torch pinned tensor .to(cuda, non_blocking)", lambda: pin.to("cuda", non_blocking=True)),
                 ("torch from_numpy(view of pinned).to(cuda)", lambda: torch.from_numpy(pin.numpy()).to("cuda")),
                 ("torch from_numpy(pageable).to(cuda)", lambda: torch.from_numpy(pag).to("cuda")),
                 ("oc_upload page-locked", lambda: ctx.upload(pin.numpy(), out)),
                 ("oc_upload pageable (staged)", lambda: ctx.upload(pag, out))):
    ms = t(fn)
    print(f"{name:48s} {ms:8.2f} ms  {gb / ms * 1e3:6.1f} GB/s", flush=True)
