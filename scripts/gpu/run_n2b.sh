set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_n2_b.err | tail -12; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_b.json')); print({k:d[k] for k in ('value','ms_per_step','parity')}); print(d['roofline']['frac'], d['e2e']['value'], d['gcfm']['value'], d['gcfm']['e2e_value'])"
timeout 900 $TR --master-port 29515 bench.py --gpus 2 --workload metro --steps 2 --warmup 1 > gpurun_out/r2_bench_metro_n2.json 2> gpurun_out/r2_bench_metro_n2.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2_bench_metro_n2.err | tail -5; cut -c1-300 gpurun_out/r2_bench_metro_n2.json
