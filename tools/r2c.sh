#!/bin/bash
# ncu of the fused step+deposit kernel (probe: 11 plain launches, then 11 fused ones); the reports are
# turned into CSV pages on the box (gpurun brings back at most 64 MiB)
set -x
mkdir -p gpurun_out
for cfg in C2 C3; do
  ncu --set full --clock-control none --import-source on -k step_kernel_wpipe -s 14 -c 1 \
      -f -o /tmp/fused_$cfg python tools/twoway_probe.py $cfg 6 > gpurun_out/r2c_ncu_$cfg.log 2>&1
  ncu -i /tmp/fused_$cfg.ncu-rep --page raw --csv > gpurun_out/r2c_fused_${cfg}_raw.csv 2>/dev/null
  ncu -i /tmp/fused_$cfg.ncu-rep --page source --csv > gpurun_out/r2c_fused_${cfg}_source.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
