#!/bin/bash
# e2e (host buffers): geometric tail of the chunk schedule vs equal chunks
mkdir -p gpurun_out
for e in 0 1; do
  if [ $e = 1 ]; then export GFSB200_HOST_EQUAL_CHUNKS=1; else unset GFSB200_HOST_EQUAL_CHUNKS; fi
  timeout 600 python tools/time_e2e.py 2>&1 | grep chunk | sed "s/^/EQUAL=$e /" | tee -a gpurun_out/r2v_e2e.log
done
