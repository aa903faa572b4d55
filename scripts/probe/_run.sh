timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; tail -2 gpurun_out/r2_bench_f.err
python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -2 gpurun_out/r2_bench_ref.err
