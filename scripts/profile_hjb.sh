#!/bin/bash
# ncu evidence for the HJB path (run under gpurun; one ncu "session" per call).
# 1) plain run must exit 0; 2) launch list with per-launch device time; 3) full capture of the top kernel;
# 4) FP64-pipe instructions per pair of the GCFM pair force (bench.py FP64_INST_PER_PAIR).
set -x
TAG=${1:-r2}
CMD="python bench.py --steps 1 --warmup 1 --gcfm-steps 2 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 || { tail -20 gpurun_out/prof_plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hjb_fused_kernel -s 6 -c 3 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
PCMD="python -m pytest tests/test_gpu_gcfm.py -q -m gpu -k test_pair_force_bitexact"
$PCMD > gpurun_out/prof_pair_plain.log 2>&1 && ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:pair_probe --csv --log-file gpurun_out/pair_fp64_$TAG.csv $PCMD > gpurun_out/ncu_pair.log 2>&1
ls -la gpurun_out/ | tail -8
