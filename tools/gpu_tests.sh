#!/bin/bash
# Runs the GPU parity tests one group at a time, each under its own hard timeout, so that a hung
# kernel in one group cannot eat the whole gpurun call.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
for t in test_k1 test_csr_rows test_csr_feature test_gram test_inverse test_fit test_admm_initialize test_admm_fixture test_admm_dense test_admm_sparse test_admm_wide test_admm_fixed_point test_admm_stop test_score test_naive_train_matches test_naive_train_many test_errors; do
  echo "=== $t"
  timeout -s KILL ${T:-240} python -m pytest tests/test_gpu_parity.py -m gpu -k $t -q -x 2>&1 | tail -${TAIL:-25}
done 2>&1 | tee gpurun_out/gpu_tests.log
echo "=== test_gpu_jobs" | tee -a gpurun_out/gpu_tests.log
timeout -s KILL ${T:-240} python -m pytest tests/test_gpu_jobs.py -m gpu -q -x 2>&1 | tail -${TAIL:-25} | tee -a gpurun_out/gpu_tests.log
