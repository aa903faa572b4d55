timeout 900 python -m pytest tests/test_gpu_extract.py -x -q 2>&1 | tail -1
timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -3
cd scripts/probe; timeout 300 python soak.py 60 7731 2>&1 | tail -1
