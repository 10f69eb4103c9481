set -x
timeout 300 python scripts/perf_batch.py 128 2.0 2>&1 | tail -3
OC_B200_LIB=$PWD/gpurun_in/liboc_b200_r2a.so timeout 300 python scripts/perf_batch.py 128 2.0 2>&1 | tail -3
timeout 300 python scripts/perf_batch.py 64 2.0 2>&1 | tail -2
