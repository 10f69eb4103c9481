set -x
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err ); echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/r2b_bench_n1.json
