set -x
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu.log
python scripts/perf_gcfm.py 12500 100000 2>&1 | tail -2
python scripts/perf_ens2.py 64 2.0 2>&1 | grep "^pass"
python scripts/perf_batch.py 64 2.0 2>&1 | tail -2
