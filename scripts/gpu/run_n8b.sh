set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29516 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_b.json 2> gpurun_out/r2_bench_n8_b.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_n8_b.err | tail -6; python scripts/show_bench.py gpurun_out/r2_bench_n8_b.json
timeout 600 $TR --master-port 29517 bench.py --gpus 8 --workload ensemble --rooms 1024 --steps 1 --warmup 1 > gpurun_out/r2_bench_ens_n8.json 2> gpurun_out/r2_bench_ens_n8.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_ens_n8.err | tail -4; python scripts/show_bench.py gpurun_out/r2_bench_ens_n8.json
