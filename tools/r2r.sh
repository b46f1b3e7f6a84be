#!/bin/bash
# timing only (A/B of library builds in ONE box): step + fused kernel
mkdir -p gpurun_out
L=$PWD/gerris-fft-particles_b200/lib
for cfg in ${CFGS:-C2}; do
  for v in ${LIBS:-default}; do
    f=$L/libgfsb200.so; [ "$v" != default ] && f=$L/libgfsb200_$v.so
    GFSB200_LIB=$f timeout 300 python tools/twoway_probe.py $cfg 40 2>&1 | tail -1 | sed "s/^/$v /" | tee -a gpurun_out/r2r_probe.log
  done
done
