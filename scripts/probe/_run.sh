timeout 900 python -m pytest tests/test_gpu_extract.py -x -q 2>&1 | tail -2
for c in kitti hd uhd; do python scripts/probe/dev_batch.py $c 20; done
NH=4 ORBX_SPLIT=1 python scripts/probe/two_handles.py kitti 20
