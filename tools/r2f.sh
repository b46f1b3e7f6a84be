#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size" > gpurun_out/r2f_fullsize.log 2>&1; tail -15 gpurun_out/r2f_fullsize.log
timeout 900 python -m pytest tests/test_dropin_module.py -m gpu -q -x > gpurun_out/r2f_module.log 2>&1; tail -5 gpurun_out/r2f_module.log
GFSB200_MODULE_PROFILE=1 timeout 600 python tools/time_module.py 10000000 > gpurun_out/r2f_time_module.log 2>&1; tail -4 gpurun_out/r2f_time_module.log
