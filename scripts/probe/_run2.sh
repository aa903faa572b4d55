timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_stereo.py tests/test_gpu_next_rows.py -x -q 2>&1 | tail -1
timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -3
cd scripts/probe; timeout 300 python soak.py 60 7725 2>&1 | tail -1; timeout 200 python soak_next.py 40 7726 2>&1 | tail -1
