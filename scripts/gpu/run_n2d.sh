set -x
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 scripts/dist_gcfm_check.py 2>&1 | grep -E "DIST_GCFM|rank|Error|error" | head
OC_RECOMPUTE=1 timeout 300 $TR --master-port 29513 scripts/dist_gcfm_check.py 2>&1 | grep -E "DIST_GCFM|Error|error" | head
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; tail -3 gpurun_out/r2b_bench_n2.err; python scripts/show_bench.py gpurun_out/r2b_bench_n2.json
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --workload metro --steps 3 --warmup 1 > gpurun_out/r2b_bench_metro_n2.json 2> gpurun_out/r2b_bench_metro_n2.err; tail -3 gpurun_out/r2b_bench_metro_n2.err; python scripts/show_bench.py gpurun_out/r2b_bench_metro_n2.json
