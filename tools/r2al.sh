#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_twoway.py -m gpu -q -x > gpurun_out/r2al_tests.log 2>&1; tail -3 gpurun_out/r2al_tests.log
timeout 300 python tools/twoway_probe.py C3 40 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/twoway_probe.py C2 40 2>&1 | tail -1 | cut -c1-200
