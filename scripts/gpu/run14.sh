set -x
timeout 900 python -m pytest tests/test_gpu_simulation.py tests/test_gpu_ensemble.py -q -m gpu -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
python scripts/perf_gcfm.py 12500 100000 2>&1 | tail -2
