set -x
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --gcfm-steps 2 --no-cpu-baseline 2>&1 >/dev/null | grep -E "device-resident|e2e"
OC_BENCH_NO_PREFETCH=1 timeout 300 python bench.py --steps 10 --warmup 3 --gcfm-steps 2 --no-cpu-baseline 2>&1 >/dev/null | grep -E "device-resident|e2e"
done
