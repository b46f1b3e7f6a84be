#!/bin/bash
mkdir -p gpurun_out
for k in 1 2 3; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-configs 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms_per_step', round(d['ms_per_step'], 5), 'value %.4g' % d['value'], d['clocks'])" | tee -a gpurun_out/r2ai.log
done
