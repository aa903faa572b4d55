set -x
python scripts/probe/soak.py 300 3027 2>&1 | tail -3
python scripts/probe/soak.py 100 3028 big 2>&1 | tail -3
python scripts/probe/soak_edges.py 2>&1 | tail -3
ORBX_FAST_KCAP=16 python scripts/probe/soak.py 120 3032 2>&1 | tail -3
python scripts/probe/soak_next.py 60 3029 2>&1 | tail -2
python scripts/probe/soak_match.py 60 3030 2>&1 | tail -2
python scripts/probe/soak_handle.py 150 3031 2>&1 | tail -2
python scripts/probe/soak_threads.py 60 2>&1 | tail -2
