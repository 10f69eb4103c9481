set -x
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 scripts/dist_check.py 2>&1 | grep -E "DIST_CHECK|Error|error" | head
timeout 300 $TR --master-port 29512 scripts/dist_gcfm_check.py 2>&1 | grep -E "DIST_GCFM|rank|Error|error" | head
OC_RECOMPUTE=1 timeout 300 $TR --master-port 29513 scripts/dist_gcfm_check.py 2>&1 | grep -E "DIST_GCFM|Error|error" | head
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err; tail -3 gpurun_out/r2_bench_n2_a.err; cat gpurun_out/r2_bench_n2_a.json | cut -c1-1500
