for v in "X=1" "X=2" "ORBX_LANES=1"; do echo "== $v"; env $v python scripts/probe/soak_handle_dbg.py 150 4031 2>&1 | grep "^call\|bad frames" | cut -c1-150 | tail -4; done
