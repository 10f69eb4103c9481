timeout 400 python -m pytest tests/test_gpu_gcfm.py tests/test_gpu_simulation.py tests/test_gpu_ensemble.py -m gpu -q -x --timeout 150 2>&1 | tail -3
timeout 200 python scripts/perf_gcfm.py 2>&1 | tail -4
timeout 200 python scripts/perf_ensemble.py 32 1.0 2>&1 | tail -3
