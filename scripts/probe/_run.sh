timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t.log 2>&1; tail -2 gpurun_out/r2_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err; tail -c 600 gpurun_out/r2_final2_bench.json
