timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t.log 2>&1; tail -2 gpurun_out/r2_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 200 gpurun_out/bench_r2.json
cd scripts/probe; for s in 4031 6031; do timeout 100 python soak_handle.py 300 $s 2>&1 | tail -1; done; cd ../..
ORBX_SPLIT=1 ncu --set full --clock-control none --import-source on --launch-skip 26 --launch-count 13 -f -o gpurun_out/prof_all_r2 python scripts/probe/one_step.py > gpurun_out/prof_all_r2.log 2>&1; tail -1 gpurun_out/prof_all_r2.log
python bench.py --quick --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2.csv python bench.py --quick --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/prof_all_r2.ncu-rep gpurun_out/launches_r2.csv
