#!/bin/bash
# module changes: whole GPU suite + the module event timing (default = resident mode)
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_gpu_tests.log 2>&1; tail -25 gpurun_out/r2e_gpu_tests.log
GFSB200_MODULE_PROFILE=1 timeout 600 python tools/time_module.py 10000000 > gpurun_out/r2e_time_module.log 2>&1; tail -5 gpurun_out/r2e_time_module.log
