set -x
timeout 900 python -m pytest tests/test_gpu_gcfm.py tests/test_gpu_simulation.py tests/test_gpu_ensemble.py -q -m gpu -x > gpurun_out/r2_pytest_gcfm.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_gcfm.log
for knobs in "" "gcfm_split=0" "gcfm_ws_pair=0"; do
OC_KNOBS=$knobs timeout 300 python scripts/perf_gcfm.py 12500 100000 2>&1 | tail -2
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|chain|cand|wall_search|agent_terms|setup_k|scan_k|scatter|noise|exit_comp" -c 120 --csv --log-file gpurun_out/launches_gcfm.csv python scripts/perf_gcfm.py 12500 > gpurun_out/ncu_launch_gcfm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|chain|cand|wall_search|agent_terms|setup_k|scan_k|scatter|noise|exit_comp" -c 70 --csv --log-file gpurun_out/launches_gcfm_100k.csv python scripts/perf_gcfm.py 100000 > gpurun_out/ncu_launch_gcfm2.log 2>&1
