set -x
timeout 900 python -m pytest tests/test_gpu_gcfm.py tests/test_gpu_simulation.py tests/test_gpu_ensemble.py -q -m gpu -x > gpurun_out/r2_pytest_gcfm.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_gcfm.log
for knobs in "" "gcfm_graph=0"; do
OC_KNOBS=$knobs timeout 300 python scripts/perf_gcfm.py 12500 100000 2>&1 | tail -2
done
