#!/bin/bash
# GPU-side capture for profiles/: tests, the bench line (both arms), the ncu launch list and one
# --set full capture of the two hot kernels.  Run under gpurun; results land in gpurun_out/.
set -x
tag=${1:-r1e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; tail -1 gpurun_out/bench_$tag.log | cut -c1-400
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_ref.log 2>&1; tail -1 gpurun_out/bench_${tag}_ref.log | cut -c1-300
python bench.py --config C3 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_${tag}_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"step_kernel|lattice_cell_pass" -s 9 -c 2 \
    -f -o gpurun_out/prof_$tag python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out | tail -8
