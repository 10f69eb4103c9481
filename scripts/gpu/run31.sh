set -x
for cfg in "0 0" "1 0" "1 4" "1 16"; do set -- $cfg
OC_BENCH_PREFETCH=$([ "$1" = 1 ] && echo 1) OC_UPLOAD_CHUNK_MB=$2 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --gcfm-steps 10 > gpurun_out/up_t.json 2> gpurun_out/up_t.err; echo "prefetch=$1 chunk=$2MB"; python scripts/show_bench.py gpurun_out/up_t.json | cut -c1-120
done
