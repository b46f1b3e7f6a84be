#!/bin/bash
# split table rows: parity first, then the timing
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_gpu_tests.log 2>&1; tail -6 gpurun_out/r2l_gpu_tests.log
for cfg in C2 C3 2D; do python tools/twoway_probe.py $cfg 40 | tee -a gpurun_out/r2l_probe.log; done
python tools/time_cellpass.py 2>&1 | tail -3 | tee gpurun_out/r2l_cellpass.log
