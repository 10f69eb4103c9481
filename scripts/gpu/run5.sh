set -x
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest_gpu.log
python scripts/perf_batch.py 64 2.0
OC_BATCH_PER_ROOM=1 python scripts/perf_batch.py 64 2.0
python scripts/perf_batch.py 128 2.0
