set -x
OC_DEBUG_TIMING=1 timeout 600 python bench.py --workload ensemble --rooms 128 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ens_t.json 2> gpurun_out/ens_t.err; grep "pass:\|oc multi" gpurun_out/ens_t.err | tail -10
OC_DEBUG_TIMING=1 OC_RNG_THREADS=0 OC_ENSEMBLE_BUILD_THREADS=1 timeout 600 python bench.py --workload ensemble --rooms 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ens_t.json 2> gpurun_out/ens_t.err; grep "pass:\|oc multi" gpurun_out/ens_t.err | tail -8
