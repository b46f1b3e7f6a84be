#!/bin/bash
# bulk-staged cell pass: bit-identity test, then timing against the gather-staged one
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lattice_cell_pass or vorticity or corner or full_size_c2" > gpurun_out/r2ab_tests.log 2>&1; tail -3 gpurun_out/r2ab_tests.log
for b in 1 0 1 0; do
  GFSB200_CELLPASS_BULK=$b python tools/time_cellpass.py C2 2>&1 | tail -1 | sed "s/^/BULK=$b /" | tee -a gpurun_out/r2ab_cellpass.log
done
