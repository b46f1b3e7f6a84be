#!/bin/bash
# round-2 evidence: the bench line at 100 steps, launch list, ncu --set full of the kernels that had none
set -x
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 10 > gpurun_out/r2i_bench.log 2> gpurun_out/r2i_bench.err; tail -c 600 gpurun_out/r2i_bench.err; tail -1 gpurun_out/r2i_bench.log | cut -c1-300
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2i_bench_ref.log 2>&1; tail -1 gpurun_out/r2i_bench_ref.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2i_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-configs > gpurun_out/r2i_ncu_launches.log 2>&1
for cfg in C3 2D; do
  ncu --set full --clock-control none --import-source on -k step_kernel_wpipe -s 6 -c 1 \
      -f -o /tmp/step_$cfg python tools/twoway_probe.py $cfg 6 > gpurun_out/r2i_ncu_step_$cfg.log 2>&1
  ncu -i /tmp/step_$cfg.ncu-rep --page raw --csv > gpurun_out/r2i_step_${cfg}_raw.csv 2>/dev/null
  ncu -i /tmp/step_$cfg.ncu-rep --page source --csv > gpurun_out/r2i_step_${cfg}_source.csv 2>/dev/null
done
ncu --set full --clock-control none --import-source on -k regex:"lattice_cell_pass" -s 4 -c 1 \
    -f -o /tmp/cellpass python tools/twoway_probe.py C2 4 > gpurun_out/r2i_ncu_cell.log 2>&1
ncu -i /tmp/cellpass.ncu-rep --page raw --csv > gpurun_out/r2i_cellpass_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
