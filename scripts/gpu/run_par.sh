set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
OC_BENCH_PARITY_ONLY=1 OC_BENCH_WATCHDOG=100 timeout 200 $TR --master-port 29514 bench.py --gpus 2 > gpurun_out/par.json 2> gpurun_out/par.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/par.err | tail -12
timeout 300 $TR --master-port 29516 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_n2_b.err | tail -8; python scripts/show_bench.py gpurun_out/r2_bench_n2_b.json
