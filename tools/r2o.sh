#!/bin/bash
# A/B of library builds (lib/libgfsb200_<MACRO>.so): contiguous tile ranges per CTA, fast special functions
mkdir -p gpurun_out
L=gerris-fft-particles_b200/lib
for cfg in C2 C3; do
  for v in "" _GFSB200_CONTIG _GFSB200_FASTMATH _GFSB200_AB_BOTH; do
    GFSB200_LIB=$PWD/$L/libgfsb200$v.so timeout 300 python tools/twoway_probe.py $cfg 40 2>&1 | tail -1 | tee -a gpurun_out/r2o_ab.log
  done
done
