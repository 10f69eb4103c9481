#!/bin/bash
# ncu evidence for the HJB path (run under gpurun; one ncu "session" per call).
# 1) plain run must exit 0; 2) launch list with per-launch device time; 3) full capture of the top kernel.
set -x
FUSED=${1:-1}
CMD="python bench.py --fused $FUSED --steps 1 --warmup 1 --gcfm-steps 2 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain_f$FUSED.log 2>&1 || { tail -20 gpurun_out/prof_plain_f$FUSED.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_f$FUSED.csv $CMD > gpurun_out/ncu_launch_f$FUSED.log 2>&1
KPAT="hjb_fused_kernel"; [ "$FUSED" = "0" ] && KPAT="hjb_stage_kernel"
ncu --set full --clock-control none --import-source on -k regex:$KPAT -s 6 -c 3 -o gpurun_out/prof_f$FUSED -f $CMD > gpurun_out/ncu_full_f$FUSED.log 2>&1
ls -la gpurun_out/
