set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_hjb.py -x -q -m gpu > gpurun_out/r2_pytest_hjb.log 2>&1; echo "pytest rc=$?" ; tail -5 gpurun_out/r2_pytest_hjb.log
python scripts/perf_fused.py 2>&1 | tail -1
for v in nobulk chain1 pf12 chain1pf12; do OC_B200_LIB=$PWD/optimal_crowds_b200/variants/liboc_$v.so python scripts/perf_fused.py 2>&1 | tail -1; done
timeout 120 scripts/probes/cdp_probe 2000
