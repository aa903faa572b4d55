timeout 900 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_next_rows.py tests/test_gpu_dropin_frame.py -x -q 2>&1 | tail -6
