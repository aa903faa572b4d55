timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_abi.py -x -q -k "pipe or abi or exported or symbol" 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -5 gpurun_out/r2_bench_c.err
