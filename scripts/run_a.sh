for v in vw3 vw3e vw2; do OC_B200_LIB=$PWD/optimal_crowds_b200/variants/liboc_$v.so python scripts/perf_fused.py 2>&1 | tail -1; done
python scripts/perf_fused.py 2>&1 | tail -1
OC_B200_LIB=$PWD/optimal_crowds_b200/variants/liboc_vw3.so python -m pytest tests/test_gpu_hjb.py -m gpu -q -x 2>&1 | tail -3
