#!/bin/bash
# micro-benchmark: vertex gather by per-lane LDG vs one TMA tensor box per 2x2x2 block of cells
mkdir -p gpurun_out
for ppc in 4.77 38.1; do
  timeout 60 tools/microbench/tma_gather_probe $ppc
done 2>&1 | tee gpurun_out/r2n_tma_gather_probe.log
