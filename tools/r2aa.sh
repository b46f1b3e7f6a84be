#!/bin/bash
mkdir -p gpurun_out
for cfg in C2 2D C3; do timeout 300 python tools/twoway_probe.py $cfg 40 2>&1 | tail -1 | cut -c1-190 | tee -a gpurun_out/r2aa_probe.log; done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2aa_bench.log 2>gpurun_out/r2aa_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aa_bench.log').read().strip().splitlines()[-1])
print('value %.4g ms %.4f'%(d['value'], d['ms_per_step']), d['roofline']['frac'])
for k,c in d['configs'].items():
    print(k, '%.4g'%c['value'], round(c['ms_per_step'],4), round(c['step_kernel_ms'],4), round(c['roofline_frac'],3), c.get('two_way',{}).get('ms_per_step'))
PY
