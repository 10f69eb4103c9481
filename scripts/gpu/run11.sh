set -x
timeout 900 python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_gcfm.py tests/test_gpu_simulation.py -q -m gpu -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest_gpu.log
python scripts/perf_ens2.py 64 2.0 2>&1 | grep "^pass"
python scripts/perf_ens2.py 128 2.0 2>&1 | grep "^pass"
