python scripts/probe/h2d_feed.py 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/probe/h2d_feed.py 2>&1 | grep "rank(s)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 scripts/probe/h2d_feed.py 2>&1 | grep "rank(s)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 scripts/probe/h2d_feed.py 2>&1 | grep "rank(s)"
