#!/bin/bash
# two GPUs: the multi-device tests (skipped on a one-GPU box), then the default bench line at N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_twoway.py tests/test_dropin_module.py -m gpu -q -x > gpurun_out/r2b_tests_n2.log 2>&1; tail -3 gpurun_out/r2b_tests_n2.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.log 2> gpurun_out/r2b_bench_n2.err; tail -c 400 gpurun_out/r2b_bench_n2.err; tail -1 gpurun_out/r2b_bench_n2.log | cut -c1-300
