timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_cpp_adapter.py -x -q -k "pipe or adapter" 2>&1 | tail -5
