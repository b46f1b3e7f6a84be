#!/bin/bash
# ncu --set full of the plain step kernel (C2) of the current build; CSV pages made on the box
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k step_kernel_wpipe -s 6 -c 1 \
    -f -o /tmp/step_C2 python tools/twoway_probe.py C2 6 > gpurun_out/r2s_ncu_step_C2.log 2>&1
ncu -i /tmp/step_C2.ncu-rep --page raw --csv > gpurun_out/r2s_step_C2_raw.csv 2>/dev/null
ncu -i /tmp/step_C2.ncu-rep --page source --csv > gpurun_out/r2s_step_C2_source.csv 2>/dev/null
ls -la gpurun_out | tail -4
