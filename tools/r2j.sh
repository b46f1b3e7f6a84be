#!/bin/bash
# A/B: programmatic dependent launch of the step kernel; occupancy-sized grid (2D); then the whole suite
set -x
mkdir -p gpurun_out
for pdl in 0 1; do
  GFSB200_PDL=$pdl python bench.py --steps 100 --warmup 10 --resort 0 --no-cpu-baseline --e2e-steps 1 --no-configs > gpurun_out/r2j_bench_pdl$pdl.log 2>&1
  tail -1 gpurun_out/r2j_bench_pdl$pdl.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PDL $pdl', d['ms_per_step'], d['step_ms'], d['roofline']['kernel_ms'], 'two-way', d['two_way']['ms_per_step'], d['two_way']['kernel_ms'])"
done
python tools/twoway_probe.py 2D 40 | tee gpurun_out/r2j_2d.log
GFSB200_WPIPE_FIXED_GRID=1 python tools/twoway_probe.py 2D 40 | tee -a gpurun_out/r2j_2d.log
python tools/twoway_probe.py C3 40 | tee -a gpurun_out/r2j_2d.log
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_gpu_tests.log 2>&1; tail -4 gpurun_out/r2j_gpu_tests.log
