set -x
for t in 8 4; do
OC_DEBUG_TIMING=1 OC_ENSEMBLE_BUILD_THREADS=$t timeout 300 python bench.py --workload ensemble --rooms 128 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ens_t.json 2> gpurun_out/ens_t.err; echo "build threads $t"; grep "pass:" gpurun_out/ens_t.err | tail -3
done
