N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29533 scripts/dist_check.py > gpurun_out/dist_check_p2p.log 2>&1; grep -n "DIST_CHECK\|Error\|error" gpurun_out/dist_check_p2p.log | head -5
timeout 300 $TR --master-port 29532 scripts/dist_gcfm_check.py > gpurun_out/dist_gcfm_p2p.log 2>&1; grep "DIST_GCFM\|rror" gpurun_out/dist_gcfm_p2p.log | head -5
timeout 400 $TR --master-port 29530 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_p2p_n$N.json 2> gpurun_out/bench_p2p_n$N.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_p2p_n$N.json').read().strip().splitlines()[-1]);print('P2P', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['share_of_step'], d['config']['parallelism'])"
OC_P2P=0 timeout 400 $TR --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_nccl_n$N.json 2> gpurun_out/bench_nccl_n$N.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_nccl_n$N.json').read().strip().splitlines()[-1]);print('NCCL', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['share_of_step'])"
