set -x
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 900 python bench.py > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err ); echo "bench rc=$?"; cat gpurun_out/r2_bench_n1_c.json
( time timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref_c.json 2> gpurun_out/r2_bench_ref_c.err ); echo "ref rc=$?"; cat gpurun_out/r2_bench_ref_c.json
