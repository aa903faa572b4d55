set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t.log 2>&1; tail -25 gpurun_out/r2_t.log
