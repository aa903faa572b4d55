set -x
timeout 900 python -m pytest tests/test_gpu_extract.py -x -q > gpurun_out/r2_t.log 2>&1; tail -5 gpurun_out/r2_t.log
for c in kitti hd uhd; do python scripts/probe/dev_batch.py $c 20; done > gpurun_out/r2_dev.txt 2>&1; cat gpurun_out/r2_dev.txt
