set -x
cd scripts/probes
./fp64_probe
for t in base swp pf12 swppf12 nobulk nosync noex nophi nostore ne0 ne1; do ./fused_probe_$t 16384 2048 20 $t; done
./fused_probe_base 16384 2048 3 && ncu --set full --clock-control none --import-source on -k regex:hjb_fused -s 3 -c 1 -o ../../gpurun_out/r2_probe_base ./fused_probe_base 16384 2048 3 > ../../gpurun_out/ncu_probe.log 2>&1
