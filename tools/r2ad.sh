#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_twoway.py -m gpu -q -x > gpurun_out/r2ad_tests.log 2>&1; tail -3 gpurun_out/r2ad_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-configs > gpurun_out/r2ad_bench.log 2>gpurun_out/r2ad_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ad_bench.log').read().strip().splitlines()[-1])
print('value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d['step_ms'], d['roofline']['frac'], d['roofline']['kernel_ms'], 'two_way', d['two_way']['ms_per_step'])
PY
