#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"lattice_cell_pass" -s 4 -c 1 \
    -f -o /tmp/cellpass python tools/twoway_probe.py C2 4 > gpurun_out/r2ac_ncu_cell.log 2>&1
ncu -i /tmp/cellpass.ncu-rep --page raw --csv > gpurun_out/r2ac_cellpass_raw.csv 2>/dev/null
ncu -i /tmp/cellpass.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2ac_cellpass_source.csv 2>/dev/null
ls -la gpurun_out | tail -3
