set -x
OC_KNOBS=gcfm_graph=0 timeout 300 python -m pytest tests/test_gpu_gcfm.py -q -m gpu -x -k "nobody_inside" 2>&1 | tail -5
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_gcfm.py -q -m gpu -x -k "nobody_inside" > gpurun_out/sanitizer_graph.log 2>&1; grep -n "Invalid\|at .*oc_gcfm\|by thread\|Address\|=========     at" gpurun_out/sanitizer_graph.log | head -30
