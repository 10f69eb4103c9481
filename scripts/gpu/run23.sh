set -x
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|MHz" | head -6
for t in 0 1 3; do OC_RNG_THREADS=$t timeout 300 python scripts/perf_gcfm.py 12500 100000 2>&1 | tail -2; done
python - <<'PY'
import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from optimal_crowds_b200 import _rng
for N in (12500, 100000):
    st = np.random.RandomState(1).get_state()
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(20):
            _rng._c_draw(st, N, N, 64)
        dt = (time.perf_counter() - t0) / 20
    print("default threads:", N, f"{dt*1e3:.3f} ms per draw")
PY
