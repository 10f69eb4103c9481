set -x
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo "bench rc=$?"; tail -4 gpurun_out/r2_bench_n1_c.err; python scripts/show_bench.py gpurun_out/r2_bench_n1_c.json
timeout 600 python bench.py --workload ensemble --rooms 128 --steps 1 --warmup 1 > gpurun_out/r2_bench_ens_c.json 2> gpurun_out/r2_bench_ens_c.err; echo "ens rc=$?"; tail -2 gpurun_out/r2_bench_ens_c.err; python scripts/show_bench.py gpurun_out/r2_bench_ens_c.json
timeout 600 python bench.py --workload metro --steps 3 --warmup 1 > gpurun_out/r2_bench_metro_c.json 2> gpurun_out/r2_bench_metro_c.err; echo "metro rc=$?"; python scripts/show_bench.py gpurun_out/r2_bench_metro_c.json
