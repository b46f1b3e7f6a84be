#!/bin/bash
# two ranks: exchange correctness (owner slices over peer memory, ncclAllReduce, no-P2P), then the
# bench at N=2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twoway.py -x -q > gpurun_out/r2d_tests.log 2>&1; tail -30 gpurun_out/r2d_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n2.log 2> gpurun_out/r2d_bench_n2.err; tail -c 2000 gpurun_out/r2d_bench_n2.err; tail -1 gpurun_out/r2d_bench_n2.log | cut -c1-400
