#!/bin/bash
# full GPU suite + the default bench line of the current build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2full_gpu_tests.log 2>&1; tail -4 gpurun_out/r2full_gpu_tests.log
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2full_bench_n1.log 2> gpurun_out/r2full_bench_n1.err; tail -c 300 gpurun_out/r2full_bench_n1.err; tail -1 gpurun_out/r2full_bench_n1.log | cut -c1-330
