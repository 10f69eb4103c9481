set -x
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1_b.err; python scripts/show_bench.py gpurun_out/r2_bench_n1_b.json
timeout 600 python bench.py --workload ensemble --rooms 128 --steps 1 --warmup 1 > gpurun_out/r2_bench_ens_b.json 2> gpurun_out/r2_bench_ens_b.err; echo "ens rc=$?"; tail -3 gpurun_out/r2_bench_ens_b.err; python scripts/show_bench.py gpurun_out/r2_bench_ens_b.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2_bench_ref.json
