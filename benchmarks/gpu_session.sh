timeout 900 python -m pytest tests/test_emat_gpu.py tests/test_reference_trace_gpu.py tests/test_free_running_gpu.py tests/test_dev_api_gpu.py -m gpu -q 2>&1 | tail -4
