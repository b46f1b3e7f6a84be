#!/bin/bash
# step-kernel time versus re-sort period (C2): does gather locality decay between sorts?
for r in 5 10 25 50 100 1000; do
  python bench.py --config ${1:-C2} --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 0 --resort $r 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('resort $r', 'ms/step %.4f kernel_ms %.4f frac %.3f sorts %d' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['sorts_in_timed_region']))"
done
