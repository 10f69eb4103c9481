#!/bin/bash
# ncu evidence for the GCFM path (run under gpurun): launch list of one perf_gcfm.py pass + full capture of the sweep.
set -x
CMD="python scripts/perf_gcfm.py"
$CMD > gpurun_out/prof_gcfm_plain.log 2>&1 || { tail -20 gpurun_out/prof_gcfm_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_gcfm.csv $CMD > gpurun_out/ncu_launch_gcfm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 8 -c 2 -o gpurun_out/prof_gcfm -f $CMD > gpurun_out/ncu_full_gcfm.log 2>&1
ls -la gpurun_out/ | grep gcfm
