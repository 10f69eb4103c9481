set -x
timeout 300 python scripts/perf_step_profile.py 12500 2>&1 | head -3
for knobs in "" "gcfm_poll_ns=100" "gcfm_poll_ns=300"; do OC_KNOBS=$knobs timeout 300 python scripts/perf_gcfm.py 12500 100000 2>&1 | tail -2; OC_KNOBS=$knobs timeout 300 python scripts/perf_dense.py 1000 2>&1 | tail -1; done
timeout 900 python bench.py --workload ensemble --rooms 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_bench_ens_n1.json 2> gpurun_out/r2b_bench_ens_n1.err; grep "pass:" gpurun_out/r2b_bench_ens_n1.err; python scripts/show_bench.py gpurun_out/r2b_bench_ens_n1.json
