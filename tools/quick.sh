#!/bin/bash
# quick A/B of the step kernel: prints value / ms_per_step / kernel_ms / frac for C2 and C3
for cfg in C2 C3; do
  python bench.py --config $cfg --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 0 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', 'value %.3e ms/step %.4f kernel_ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
