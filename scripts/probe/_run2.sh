timeout 600 python -m pytest tests/test_gpu_extract.py -x -q 2>&1 | tail -1
for rep in 1 2; do
echo "--- fused, hint 9"; timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -2
echo "--- fused, hint 10"; ORBX_LIB=$PWD/scripts/probe/_libs/liborbx_m10.so timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -2
done
cd scripts/probe; timeout 300 python soak.py 80 7724 2>&1 | tail -1
