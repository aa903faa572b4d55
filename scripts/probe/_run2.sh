timeout 600 python -m pytest tests/test_gpu_match.py tests/test_gpu_multi.py -x -q 2>&1 | tail -3
for rep in 1 2; do
timeout 200 python scripts/probe/sharded_one.py 2>&1 | tail -3; timeout 100 python scripts/probe/match_rate.py 2>&1 | tail -1
done
timeout 100 python scripts/probe/soak_match.py 100 7712 2>&1 | tail -1
