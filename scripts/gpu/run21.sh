set -x
( time timeout 900 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err ); echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/r2b_bench_n1.json 2>/dev/null || cat gpurun_out/r2b_bench_n1.json | cut -c1-3000
( time timeout 900 python bench.py --workload ensemble --rooms 128 > gpurun_out/r2b_bench_ens_n1.json 2> gpurun_out/r2b_bench_ens_n1.err ); echo "ens rc=$?"; cat gpurun_out/r2b_bench_ens_n1.json | cut -c1-3000
( time timeout 900 python bench.py --workload metro > gpurun_out/r2b_bench_metro_n1.json 2> gpurun_out/r2b_bench_metro_n1.err ); echo "metro rc=$?"; cat gpurun_out/r2b_bench_metro_n1.json | cut -c1-3000
