set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; echo "bench rc=$?"; tail -8 gpurun_out/r2_bench_n1_a.err; cut -c1-600 gpurun_out/r2_bench_n1_a.json
timeout 600 python bench.py --workload ensemble --rooms 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_ens_a.json 2> gpurun_out/r2_bench_ens_a.err; echo "ens rc=$?"; tail -5 gpurun_out/r2_bench_ens_a.err; cut -c1-1500 gpurun_out/r2_bench_ens_a.json
timeout 600 python bench.py --workload metro --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_metro_a.json 2> gpurun_out/r2_bench_metro_a.err; echo "metro rc=$?"; tail -5 gpurun_out/r2_bench_metro_a.err; cut -c1-1500 gpurun_out/r2_bench_metro_a.json
