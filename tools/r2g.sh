#!/bin/bash
# N = 8 (or $1): two-way correctness, exchange probe in both modes, the driver's bench line
N=${1:-8}
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $T --master-port 29541 tools/check_twoway_ranks.py > gpurun_out/r2g_check_n$N.log 2>&1; tail -3 gpurun_out/r2g_check_n$N.log
timeout 600 $T --master-port 29542 tools/exchange_probe.py > gpurun_out/r2g_probe_n$N.log 2>&1; tail -1 gpurun_out/r2g_probe_n$N.log
GFSB200_EXCHANGE=1 timeout 600 $T --master-port 29543 tools/exchange_probe.py > gpurun_out/r2g_probe_allreduce_n$N.log 2>&1; tail -1 gpurun_out/r2g_probe_allreduce_n$N.log
timeout 900 $T --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2g_bench_n$N.log 2> gpurun_out/r2g_bench_n$N.err; tail -c 1500 gpurun_out/r2g_bench_n$N.err; tail -1 gpurun_out/r2g_bench_n$N.log | cut -c1-300
