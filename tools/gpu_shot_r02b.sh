#!/bin/bash
# Round-2 verification of the TF32 merges on ONE B200 (gpurun): GPU parity suite and bench with MLEASE_MERGE_TF32=1, the
# factorisation of one 10k-wide system timed both ways, and the launch list of the factorisation kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MLEASE_MERGE_TF32=1
(timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_gputests.log)
tail -4 gpurun_out/r02b_gputests.log
timeout 60 python bench.py --steps 20 --warmup 3 --also "" --no-cpu > gpurun_out/r02b_bench_tf32.json 2> gpurun_out/r02b_bench_tf32.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r02b_bench_tf32.json").read().strip().splitlines()[-1])
    print("value", d["value"], "kernel_ms", d["kernel_ms"], "solver", d["solver"], "parity", d.get("parity", {}).get("rel_err_z_per_lambda"), "e2e", d["e2e"]["value"])
except Exception as e:
    print("no bench line", e)
P
(MLEASE_MERGE_TF32=0 WHICH=cholesky ROWS=100000 timeout 40 python tools/time_gram.py; MLEASE_MERGE_TF32=1 WHICH=cholesky ROWS=100000 timeout 40 python tools/time_gram.py) 2>&1 | grep cholesky | tee gpurun_out/r02b_cholesky_ms.txt
WHICH=cholesky ROWS=100000 timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'merge_tf32|dgemm_kernel|trinv|ysym' -c 400 --csv \
  --log-file gpurun_out/r02b_chol_launches_tf32.csv python tools/time_gram.py > /dev/null 2>&1
echo "ncu rc=$?"
