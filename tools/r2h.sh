#!/bin/bash
# two GPUs: module in GFSB200_DEVICES=2 mode, two-way tests, scope experiment on the C5 tree
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin_module.py -m gpu -q -x -k "two_devices" > gpurun_out/r2h_module2.log 2>&1; tail -15 gpurun_out/r2h_module2.log
timeout 900 python -m pytest tests/test_gpu_twoway.py -m gpu -q -x > gpurun_out/r2h_twoway.log 2>&1; tail -5 gpurun_out/r2h_twoway.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29551 tools/exchange_probe.py 25000000 C5 > gpurun_out/r2h_probe_c5_sys.log 2>&1; tail -1 gpurun_out/r2h_probe_c5_sys.log
GFSB200_LOCAL_RED_GPU_SCOPE=1 timeout 600 $T --master-port 29552 tools/exchange_probe.py 25000000 C5 > gpurun_out/r2h_probe_c5_gpu.log 2>&1; tail -1 gpurun_out/r2h_probe_c5_gpu.log
timeout 600 $T --master-port 29553 tools/exchange_probe.py 10000000 C2 > gpurun_out/r2h_probe_c2.log 2>&1; tail -1 gpurun_out/r2h_probe_c2.log
