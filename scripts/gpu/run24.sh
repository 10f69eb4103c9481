set -x
timeout 300 python scripts/perf_step_profile.py 12500 2>&1 | head -40
for knobs in "" "gcfm_split=0,gcfm_margin_mm=1000,gcfm_fov_cull=0,gcfm_graph=0,gcfm_overlap=0"; do OC_KNOBS=$knobs timeout 300 python scripts/perf_dense.py 1000 2>&1 | tail -1; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|chain|cand|wall_search|agent_terms|setup_k|scan_k|scatter|noise|exit_comp" -c 130 --csv --log-file gpurun_out/r2b_gcfm_launches.csv python scripts/perf_gcfm.py 12500 > gpurun_out/ncu_launch_gcfm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 8 -c 2 -o gpurun_out/prof_gcfm_chain -f python scripts/perf_gcfm.py 12500 > gpurun_out/ncu_full_chain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wall_search_kernel|cand_kernel" -s 16 -c 2 -o gpurun_out/prof_gcfm_pre -f python scripts/perf_gcfm.py 12500 > gpurun_out/ncu_full_pre.log 2>&1
ls -la gpurun_out/*.ncu-rep
