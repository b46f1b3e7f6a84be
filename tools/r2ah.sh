#!/bin/bash
mkdir -p gpurun_out
L=$PWD/gerris-fft-particles_b200/lib
for v in default CELL5 default CELL5; do
  f=$L/libgfsb200.so; [ "$v" != default ] && f=$L/libgfsb200_$v.so
  GFSB200_LIB=$f python tools/time_cellpass.py C2 2>&1 | tail -1 | sed "s/^/$v /" | tee -a gpurun_out/r2ah_cellpass.log
done
