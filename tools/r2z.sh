#!/bin/bash
# force-recording step kernel (GModule): this build vs the build at the start of the session; gather prefetch distance
mkdir -p gpurun_out
L=$PWD/gerris-fft-particles_b200/lib
for v in default r2start default r2start; do
  f=$L/libgfsb200.so; [ "$v" != default ] && f=$L/libgfsb200_$v.so
  GFSB200_LIB=$f timeout 300 python tools/rec_probe.py C2 20 2>&1 | tail -1 | tee -a gpurun_out/r2z_rec.log
done
for pf in 64 128 256; do
  GFSB200_GATHER_PF=$pf GFSB200_MODULE_PROFILE=1 timeout 600 python tools/time_module.py 10000000 5 2>&1 | grep -E "^module|particulates" | sed "s/^/PF=$pf /" | tee -a gpurun_out/r2z_module.log
done
