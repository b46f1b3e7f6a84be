#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twoway.py -x -q > gpurun_out/r2m_tests.log 2>&1; tail -5 gpurun_out/r2m_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2m_bench_n2.log 2> gpurun_out/r2m_bench_n2.err; tail -c 800 gpurun_out/r2m_bench_n2.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_n1.log 2> gpurun_out/r2m_bench_n1.err; tail -c 500 gpurun_out/r2m_bench_n1.err
