timeout 1200 python -m pytest tests/test_gpu_extract.py tests/test_gpu_stereo.py -x -q 2>&1 | tail -3
for sd in 4031 5031; do python scripts/probe/soak_handle.py 150 $sd 2>&1 | tail -1; done
for c in kitti hd uhd; do python scripts/probe/dev_batch.py $c 20; done
NH=4 ORBX_SPLIT=1 python scripts/probe/two_handles.py kitti 20
