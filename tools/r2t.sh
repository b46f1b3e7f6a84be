#!/bin/bash
# TMA box gathers: parity first, then A/B against the per-lane gathers (GFSB200_TMA_GATHER=0) in one box
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_twoway.py -m gpu -q -x > gpurun_out/r2t_tests.log 2>&1; tail -4 gpurun_out/r2t_tests.log
for rep in 1 2; do
  for g in 1 0; do
    GFSB200_TMA_GATHER=$g timeout 300 python tools/twoway_probe.py C2 40 2>&1 | tail -1 | sed "s/^/TMA_GATHER=$g /" | tee -a gpurun_out/r2t_probe.log
  done
done
