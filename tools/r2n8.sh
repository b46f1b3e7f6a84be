#!/bin/bash
# eight GPUs: the default bench line (what the driver's scaling run launches), then the rank check of the two-way exchange
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n8.log 2> gpurun_out/r2b_bench_n8.err; tail -c 400 gpurun_out/r2b_bench_n8.err; tail -1 gpurun_out/r2b_bench_n8.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 \
   tools/check_twoway_ranks.py > gpurun_out/r2b_check_n8.log 2>&1; tail -2 gpurun_out/r2b_check_n8.log | cut -c1-300
