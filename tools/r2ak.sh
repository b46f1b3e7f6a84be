#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_twoway.py -m gpu -q -x > gpurun_out/r2ak_tests.log 2>&1; tail -4 gpurun_out/r2ak_tests.log
