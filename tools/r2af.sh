#!/bin/bash
mkdir -p gpurun_out
for smp in 1 0 4 1 0 4; do
GFSB200_TIMER_SAMPLE=$smp python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-configs 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('SAMPLE=$smp', 'ms_per_step', round(d['ms_per_step'], 5), 'value %.4g' % d['value'], d['step_ms'], 'kernel_ms', d['roofline']['kernel_ms'], d['roofline']['kernel_launches'], 'two_way', round(d['two_way']['ms_per_step'],5))" | tee -a gpurun_out/r2af.log
done
