for rep in 1 2; do
echo "--- baseline (10)"; timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -2
echo "--- 9"; ORBX_LIB=$PWD/scripts/probe/_libs/liborbx_f9.so timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -2
echo "--- 8"; ORBX_LIB=$PWD/scripts/probe/_libs/liborbx_f8.so timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -2
done
timeout 600 python -m pytest tests/test_gpu_extract.py -x -q 2>&1 | tail -2
cd scripts/probe; timeout 300 python soak.py 100 7720 2>&1 | tail -1
