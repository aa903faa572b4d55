timeout 600 python -m pytest tests/test_gpu_extract.py -x -q 2>&1 | tail -2
for rep in 1 2; do
echo "--- TMA zero"; timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -2
echo "--- thread zero"; ORBX_LIB=$PWD/scripts/probe/_libs/liborbx_zt.so timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -2
done
cd scripts/probe; timeout 300 python soak.py 100 7721 2>&1 | tail -1; timeout 200 python soak_edges.py 2>&1 | tail -1
