#!/bin/bash
# lean lattice locate: parity, then timing (state hash must equal the previous build's: 580b06168021 on C2)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_twoway.py -m gpu -q -x > gpurun_out/r2p_tests.log 2>&1; tail -4 gpurun_out/r2p_tests.log
for cfg in C2 2D C3; do timeout 300 python tools/twoway_probe.py $cfg 40 2>&1 | tail -1 | tee -a gpurun_out/r2p_probe.log; done
