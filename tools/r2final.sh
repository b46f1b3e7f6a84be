#!/bin/bash
# final capture of round 2 (second session): full GPU suite, smoke, both bench arms, ncu launch list of the
# bench command, ncu --set full of the step kernel, the fused kernel and the cell pass (CSV pages made here)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_gpu_tests.log 2>&1; tail -3 gpurun_out/r2b_gpu_tests.log
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2b_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.log 2> gpurun_out/r2b_bench_ref.err; tail -1 gpurun_out/r2b_bench_ref.log | cut -c1-300
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.log 2> gpurun_out/r2b_bench_n1.err; tail -1 gpurun_out/r2b_bench_n1.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-configs > gpurun_out/r2b_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k step_kernel_wpipe -s 6 -c 1 \
    -f -o /tmp/step_C2 python tools/twoway_probe.py C2 6 > gpurun_out/r2b_ncu_step_C2.log 2>&1
ncu -i /tmp/step_C2.ncu-rep --page raw --csv > gpurun_out/r2b_step_C2_raw.csv 2>/dev/null
ncu -i /tmp/step_C2.ncu-rep --page source --csv > gpurun_out/r2b_step_C2_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k step_kernel_wpipe -s 14 -c 1 \
    -f -o /tmp/fused_C2 python tools/twoway_probe.py C2 6 > gpurun_out/r2b_ncu_fused_C2.log 2>&1
ncu -i /tmp/fused_C2.ncu-rep --page raw --csv > gpurun_out/r2b_fused_C2_raw.csv 2>/dev/null
ncu -i /tmp/fused_C2.ncu-rep --page source --csv > gpurun_out/r2b_fused_C2_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"lattice_cell_pass" -s 4 -c 1 \
    -f -o /tmp/cellpass python tools/twoway_probe.py C2 4 > gpurun_out/r2b_ncu_cell.log 2>&1
ncu -i /tmp/cellpass.ncu-rep --page raw --csv > gpurun_out/r2b_cellpass_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
