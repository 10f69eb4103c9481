timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_gcfm_check.py > gpurun_out/dist_gcfm.log 2>&1
grep -n "DIST_GCFM\|rank 0\|rank 1\|Error" gpurun_out/dist_gcfm.log | head -30
OC_RECOMPUTE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_gcfm_check.py > gpurun_out/dist_gcfm_rc.log 2>&1
grep -n "DIST_GCFM\|rank 0\|rank 1\|Error" gpurun_out/dist_gcfm_rc.log | head -30
