set -x
timeout 900 python -m pytest tests/test_gpu_gcfm.py -q -m gpu -x > gpurun_out/r2_pytest_gcfm.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gcfm.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|wall_search|agent_terms|setup_k|scan_k|scatter|noise|exit_comp" -c 120 --csv --log-file gpurun_out/launches_gcfm.csv python scripts/perf_gcfm.py 12500 > gpurun_out/ncu_launch_gcfm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|wall_search|agent_terms|setup_k|scan_k|scatter|noise|exit_comp" -c 60 --csv --log-file gpurun_out/launches_gcfm_100k.csv python scripts/perf_gcfm.py 100000 > gpurun_out/ncu_launch_gcfm2.log 2>&1
