timeout 900 python -m pytest tests/test_gpu_stereo.py -x -q 2>&1 | tail -5
