timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t.log 2>&1; tail -2 gpurun_out/r2_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2>/dev/null; tail -c 300 gpurun_out/r2_bench_ref.json
