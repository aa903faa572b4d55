timeout 900 python -m pytest tests/test_gpu_drivers.py -x -q -s 2>&1 | tail -25
