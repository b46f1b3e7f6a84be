#!/bin/bash
# final single-GPU lines of the round: full GPU suite, smoke, both bench arms (default command lines)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_gpu_tests.log 2>&1; tail -3 gpurun_out/r2b_gpu_tests.log
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2b_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.log 2> gpurun_out/r2b_bench_ref.err; tail -1 gpurun_out/r2b_bench_ref.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.log 2> gpurun_out/r2b_bench_n1.err; tail -1 gpurun_out/r2b_bench_n1.log | cut -c1-300
python bench.py > gpurun_out/r2b_bench_n1_default.log 2> gpurun_out/r2b_bench_n1_default.err; tail -1 gpurun_out/r2b_bench_n1_default.log | cut -c1-300
