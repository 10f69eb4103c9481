#!/bin/bash
# ncu evidence for the GCFM path (run under gpurun).
set -x
CMD="python scripts/perf_gcfm.py"
$CMD > gpurun_out/prof_gcfm_plain.log 2>&1 || { tail -20 gpurun_out/prof_gcfm_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 20 -c 2 -o gpurun_out/prof_gcfm -f $CMD > gpurun_out/ncu_full_gcfm.log 2>&1
ls -la gpurun_out/ | grep gcfm
