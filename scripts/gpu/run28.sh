set -x
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"multi_kernel" -s 220 -c 110 --csv --log-file gpurun_out/r2b_ens_launches.csv python bench.py --workload ensemble --rooms 128 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_ens.log 2>&1
tail -3 gpurun_out/ncu_ens.log
