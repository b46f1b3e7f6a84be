#!/bin/bash
# NODATA interpolation out of line: new parity test + the parity files, then timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_twoway.py -m gpu -q -x > gpurun_out/r2q_tests.log 2>&1; tail -4 gpurun_out/r2q_tests.log
for cfg in C2 C3 2D; do timeout 300 python tools/twoway_probe.py $cfg 40 2>&1 | tail -1 | tee -a gpurun_out/r2q_probe.log; done
