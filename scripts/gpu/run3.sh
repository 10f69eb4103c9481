set -x
timeout 600 python -m pytest tests/test_gpu_hjb.py -x -q -m gpu > gpurun_out/r2_pytest_hjb.log 2>&1; echo "pytest rc=$?" ; tail -5 gpurun_out/r2_pytest_hjb.log
python scripts/perf_fused.py 2>&1 | tail -1
cd scripts/probes
for t in base nobulk nophi nostore ne0 ne1 ne2; do ./fused_probe_$t 16384 2048 20 $t; done
./fused_probe_base 16384 2048 3 && ncu --set full --clock-control none --import-source on -k regex:hjb_fused -s 3 -c 1 -o ../../gpurun_out/r2_probe_tma ./fused_probe_base 16384 2048 3 > ../../gpurun_out/ncu_probe.log 2>&1
