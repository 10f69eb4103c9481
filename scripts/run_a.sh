N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29533 scripts/dist_check.py > gpurun_out/dist_check_p2p_n$N.log 2>&1; grep -n "DIST_CHECK" gpurun_out/dist_check_p2p_n$N.log | head -3
timeout 120 python -m pytest tests/test_gpu_hjb.py -m gpu -q -x --timeout 100 2>&1 | tail -2
