#!/bin/bash
# ncu captures for profiles/: launch list (share of step) + full-set capture of the two top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $B --steps 3 --warmup 1 > gpurun_out/ncu_launch_bench.log 2>&1
# K1: skip the first 8 launches (cold-start iteration incl. its emit passes) and capture two steady-state passes
ncu --set full --clock-control none --import-source on -k regex:k1_dense -s 8 -c 2 -o gpurun_out/prof_k1 -f $B --steps 4 --warmup 0 > gpurun_out/ncu_k1.log 2>&1
# K1 emit pass (first launch) and the Gram kernel
ncu --set full --clock-control none --import-source on -k regex:k1_dense -c 1 -o gpurun_out/prof_k1_emit -f $B --steps 1 --warmup 0 > gpurun_out/ncu_k1e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_tcgen05 -c 1 -o gpurun_out/prof_gram -f $B --steps 1 --warmup 0 > gpurun_out/ncu_gram.log 2>&1
ls -la gpurun_out | head -30
