timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err
tail -c 400 gpurun_out/r2_bench_n8b.json; tail -3 gpurun_out/r2_bench_n8b.err
