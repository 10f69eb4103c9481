timeout 500 python -m pytest tests/test_gpu_gcfm.py tests/test_gpu_simulation.py -m gpu -q -x --timeout 200 2>&1 | tail -3
timeout 200 python scripts/perf_gcfm.py 2>&1 | tail -4
