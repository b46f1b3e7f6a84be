#!/bin/bash
# A/B: one-tile lookahead (locate the next tile, prefetch its table rows into L1)
set -x
mkdir -p gpurun_out
for cfg in C2 C3 2D; do
  for la in 0 1; do
    GFSB200_LOOKAHEAD=$la python tools/twoway_probe.py $cfg 40 | sed "s/^/LOOKAHEAD=$la /" | tee -a gpurun_out/r2k_lookahead.log
  done
done
