timeout 500 python -m pytest tests/test_gpu_hjb.py::test_full_size_band_properties tests/test_gpu_gcfm.py::test_sweep_result_is_independent_of_the_schedule -m gpu -q -x --timeout 240 2>&1 | tail -12
