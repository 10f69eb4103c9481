set -x
timeout 600 python -m pytest tests/test_gpu_ensemble.py -q -m gpu -x 2>&1 | tail -2
timeout 600 python bench.py --workload ensemble --rooms 128 --steps 3 --warmup 1 > gpurun_out/r2b_bench_ens_n1.json 2> gpurun_out/r2b_bench_ens_n1.err; grep "pass:" gpurun_out/r2b_bench_ens_n1.err | tail -3; python scripts/show_bench.py gpurun_out/r2b_bench_ens_n1.json
