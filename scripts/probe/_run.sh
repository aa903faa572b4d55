timeout 900 python -m pytest tests/test_gpu_extract.py -x -q 2>&1 | tail -3
ORBX_FAST_KCAP=8 timeout 900 python -m pytest tests/test_gpu_extract.py -x -q 2>&1 | tail -3
for c in kitti hd uhd; do python scripts/probe/dev_batch.py $c 20; done
