# extended soak with fresh seeds (not part of the test suite)
cd scripts/probe
{
timeout 400 python soak.py 500 7701
timeout 300 python soak.py 40 7702 big
timeout 200 python soak_edges.py
timeout 200 python soak_next.py 100 7703
timeout 200 python soak_match.py 100 7704
for s in 9031 9032 9033 9034 9035 9036; do timeout 100 python soak_handle.py 300 $s; done
timeout 100 python soak_threads.py 100
} > ../../gpurun_out/soak_ext.log 2>&1
grep -c . ../../gpurun_out/soak_ext.log
grep -i "soak\|bad\|MISMATCH\|error" ../../gpurun_out/soak_ext.log | tail -30
