set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29516 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_n2_c.err | tail -4; python scripts/show_bench.py gpurun_out/r2_bench_n2_c.json
OC_STEPS=40 timeout 200 $TR --master-port 29512 scripts/dist_gcfm_check.py 2>&1 | grep -E "DIST_GCFM|Error|error" | head
