#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus ${NG:-2} --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2aj_bench.log 2> gpurun_out/r2aj_bench.err; tail -c 600 gpurun_out/r2aj_bench.err; tail -1 gpurun_out/r2aj_bench.log | cut -c1-200
