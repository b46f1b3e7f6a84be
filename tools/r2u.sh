#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench_n1.log 2> gpurun_out/r2u_bench_n1.err; tail -c 300 gpurun_out/r2u_bench_n1.err; tail -1 gpurun_out/r2u_bench_n1.log | cut -c1-400
