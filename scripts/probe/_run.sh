set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1
tail -15 gpurun_out/r2_gputests.log
