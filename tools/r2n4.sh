#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n4.log 2> gpurun_out/r2b_bench_n4.err; tail -c 300 gpurun_out/r2b_bench_n4.err; tail -1 gpurun_out/r2b_bench_n4.log | cut -c1-300
