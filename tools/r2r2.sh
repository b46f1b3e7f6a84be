#!/bin/bash
# kernel A/B on three configs + whole-step (bench, one-way and two-way) for each library build
mkdir -p gpurun_out
L=$PWD/gerris-fft-particles_b200/lib
for v in ${LIBS:-default}; do
  f=$L/libgfsb200.so; [ "$v" != default ] && f=$L/libgfsb200_$v.so
  for cfg in C2 C3 2D; do
    GFSB200_LIB=$f timeout 300 python tools/twoway_probe.py $cfg 40 2>&1 | tail -1 | sed "s/^/$v /" | tee -a gpurun_out/r2r2_probe.log
  done
  GFSB200_LIB=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$v', 'bench ms_per_step', round(d['ms_per_step'], 5), 'value %.4g' % d['value'], 'kernel_ms', round(d['roofline']['kernel_ms'], 5), 'two_way ms', round(d['two_way']['ms_per_step'], 5), 'e2e_res %.4g' % d['e2e_resident']['value'])" | tee -a gpurun_out/r2r2_probe.log
done
