#!/bin/bash
# round-2 baseline: two-way at N=1 (deposit kernel time), before any change
set -x
mkdir -p gpurun_out
python bench.py --two-way --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2a_twoway_c2.log 2>&1; tail -1 gpurun_out/r2a_twoway_c2.log | cut -c1-300
python bench.py --config C3 --two-way --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2a_twoway_c3.log 2>&1; tail -1 gpurun_out/r2a_twoway_c3.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2a_launches_twoway.csv \
    python bench.py --two-way --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2a_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"deposit_kernel" -s 4 -c 1 \
    -f -o gpurun_out/r2a_deposit python bench.py --two-way --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2a_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
