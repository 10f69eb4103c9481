set -x
timeout 600 python -m pytest tests/test_gpu_ensemble.py -q -m gpu -x 2>&1 | tail -3
for cfg in "0 " "2 " "2 gcfm_poll_ns=400" "4 gcfm_poll_ns=400" "2 gcfm_poll_ns=1000"; do set -- $cfg
OC_ENSEMBLE_CTAS=$1 OC_KNOBS=$2 timeout 900 python bench.py --workload ensemble --rooms 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ens_t.json 2> gpurun_out/ens_t.err; echo "cfg $cfg"; grep "pass:" gpurun_out/ens_t.err | tail -1; python scripts/show_bench.py gpurun_out/ens_t.json
done
