#!/bin/bash
# round-2 validation: new two-way tests first, then the whole GPU suite, then a short bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_twoway.py -x -q > gpurun_out/r2b_twoway_tests.log 2>&1; tail -25 gpurun_out/r2b_twoway_tests.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_gpu_tests.log 2>&1; tail -15 gpurun_out/r2b_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.log 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.err; tail -1 gpurun_out/r2b_bench.log | cut -c1-600
