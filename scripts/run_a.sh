N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29530 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.json | cut -c1-900
if [ "$N" = "8" ]; then
timeout 400 $TR --master-port 29531 scripts/dist_room.py > gpurun_out/dist_room_n$N.log 2>&1; grep "what" gpurun_out/dist_room_n$N.log
timeout 300 $TR --master-port 29532 scripts/dist_gcfm_check.py > gpurun_out/dist_gcfm_n$N.log 2>&1; grep "DIST_GCFM" gpurun_out/dist_gcfm_n$N.log
timeout 300 $TR --master-port 29533 scripts/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1; grep "DIST_CHECK" gpurun_out/dist_check_n$N.log
fi
