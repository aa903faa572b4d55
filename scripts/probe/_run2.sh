for rep in 1 2; do
echo "--- 8"; timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -2
for v in p6 p12 p16; do echo "--- $v"; ORBX_LIB=$PWD/scripts/probe/_libs/liborbx_$v.so timeout 200 python scripts/probe/stage_times64.py 2>&1 | tail -2; done
done
