timeout 700 python -m pytest tests -m gpu -q --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; grep "bench " gpurun_out/bench_n1.err | tail -4
timeout 500 bash scripts/profile_hjb.sh 1 2>&1 | tail -2
timeout 500 bash scripts/profile_gcfm.sh 2>&1 | tail -2
