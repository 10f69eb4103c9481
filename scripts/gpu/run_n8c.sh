set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29516 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2b_bench_n8.json 2> gpurun_out/r2b_bench_n8.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2b_bench_n8.err | tail -5; python scripts/show_bench.py gpurun_out/r2b_bench_n8.json
timeout 300 $TR --master-port 29517 bench.py --gpus 8 --workload ensemble --rooms 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_bench_ens_n8.json 2> gpurun_out/r2b_bench_ens_n8.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2b_bench_ens_n8.err | tail -3; python scripts/show_bench.py gpurun_out/r2b_bench_ens_n8.json
