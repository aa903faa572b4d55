timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
tail -3 gpurun_out/r2_bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n4.json 2> gpurun_out/r2_bench_ref_n4.err
tail -2 gpurun_out/r2_bench_ref_n4.err; wc -c gpurun_out/r2_bench_ref_n4.json
