#!/bin/bash
mkdir -p gpurun_out
L=$PWD/gerris-fft-particles_b200/lib
for v in default GFSB200_V_NOHINT default GFSB200_V_NOHINT; do
  f=$L/libgfsb200.so; [ "$v" != default ] && f=$L/libgfsb200_$v.so
  GFSB200_LIB=$f timeout 300 python tools/twoway_probe.py C2 40 2>&1 | tail -1 | sed "s/^/$v /" | cut -c1-200 | tee -a gpurun_out/r2ae_probe.log
done
ncu --set full --clock-control none --import-source on -k regex:"lattice_cell_pass" -s 4 -c 1 \
    -f -o /tmp/cellpass python tools/twoway_probe.py C2 4 > gpurun_out/r2ae_ncu_cell.log 2>&1
ncu -i /tmp/cellpass.ncu-rep --page raw --csv > gpurun_out/r2ae_cellpass_raw.csv 2>/dev/null
ncu -i /tmp/cellpass.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2ae_cellpass_source.csv 2>/dev/null
