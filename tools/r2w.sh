#!/bin/bash
# ncu --set full of the fused step+deposit kernel (C2) of the current build
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k step_kernel_wpipe -s 14 -c 1 \
    -f -o /tmp/fused_C2 python tools/twoway_probe.py C2 6 > gpurun_out/r2w_ncu_fused_C2.log 2>&1
ncu -i /tmp/fused_C2.ncu-rep --page raw --csv > gpurun_out/r2w_fused_C2_raw.csv 2>/dev/null
ncu -i /tmp/fused_C2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2w_fused_C2_source.csv 2>/dev/null
ls -la gpurun_out | tail -3
