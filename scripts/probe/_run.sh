timeout 900 python -m pytest tests/test_cpp_adapter.py tests/test_gpu_dropin_frame.py tests/test_gpu_drivers.py tests/test_gpu_dropin_vocabulary.py tests/test_abi.py -x -q 2>&1 | tail -4
python scripts/probe/frame_constructor_times.py
