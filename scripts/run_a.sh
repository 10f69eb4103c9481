for v in lcc2 c3; do OC_B200_LIB=$PWD/optimal_crowds_b200/variants/liboc_$v.so python scripts/perf_fused.py 2>&1 | tail -1; done
python scripts/perf_fused.py 2>&1 | tail -1
bash scripts/profile_hjb.sh 1 2>&1 | tail -3
