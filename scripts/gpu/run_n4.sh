set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29518 bench.py --gpus 4 --workload metro --steps 5 --warmup 2 > gpurun_out/r2_bench_metro_n4.json 2> gpurun_out/r2_bench_metro_n4.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_metro_n4.err | tail -4; python scripts/show_bench.py gpurun_out/r2_bench_metro_n4.json
timeout 420 $TR --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_n4.err | tail -3; python scripts/show_bench.py gpurun_out/r2_bench_n4.json
