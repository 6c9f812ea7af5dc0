#!/bin/bash
# ncu captures for profiles/: launch list (share of step) + full-set capture of the two top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu"
MLEASE_DEBUG=1 $B --steps 12 --warmup 0 > gpurun_out/trace.json 2> gpurun_out/trace.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B --steps 2 --warmup 1 > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_dense -s 2 -c 2 -o gpurun_out/prof_k1 -f $B --steps 1 --warmup 0 > gpurun_out/ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_tcgen05 -c 1 -o gpurun_out/prof_gram -f $B --steps 1 --warmup 0 > gpurun_out/ncu_gram.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/ncu_k1.log
