timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_abi.py -x -q -k "pipe or abi or exported or symbol or device" 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err
