set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29516 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_a.json 2> gpurun_out/r2_bench_n8_a.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_n8_a.err | tail -8; python scripts/show_bench.py gpurun_out/r2_bench_n8_a.json
