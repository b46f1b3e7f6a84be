#!/bin/bash
# the GModule's field gather: prefetch distance and streaming stores
mkdir -p gpurun_out
for pf in 32 64; do for nt in 0 1; do
  GFSB200_GATHER_PF=$pf GFSB200_GATHER_NT=$nt GFSB200_MODULE_PROFILE=1 timeout 600 python tools/time_module.py 10000000 5 2>&1 | grep -E "^module|particulates" | sed "s/^/PF=$pf NT=$nt /" | tee -a gpurun_out/r2y_module.log
done; done
