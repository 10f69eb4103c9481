set -x
timeout 900 python -m pytest tests/test_gpu_simulation.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python scripts/perf_advance.py 12500 100000 2>&1 | tail -2
OC_RNG_THREADS=0 timeout 300 python scripts/perf_advance.py 12500 100000 2>&1 | tail -2
