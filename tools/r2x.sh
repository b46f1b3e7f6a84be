#!/bin/bash
# parity (both GPU parity files), then step + fused kernel timing on C2 / C3 / 2D
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_twoway.py -m gpu -q -x > gpurun_out/r2x_tests.log 2>&1; tail -3 gpurun_out/r2x_tests.log
for cfg in ${CFGS:-C2 C3 2D}; do timeout 300 python tools/twoway_probe.py $cfg 40 2>&1 | tail -1 | tee -a gpurun_out/r2x_probe.log; done
