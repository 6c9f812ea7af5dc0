#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun): GPU tests, the default bench line, the launch list of the same command restricted
# to the library's kernels, and one `ncu --set full` capture of each roofline kernel.  Summaries are made HERE afterwards with
# tools/summarize_r02.py (ncu / list / traffic / sass) and committed under profiles/; gpurun_out/ is scratch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r02}
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
Q="--also '' --no-cpu --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mlease -c 4000 --csv \
  --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --also "" --no-cpu --no-parity > /dev/null 2> gpurun_out/${T}_ncu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_csr -s 1 -c 1 -o gpurun_out/${T}_gram_csr -f \
  python bench.py --steps 2 --warmup 1 --also "" --no-cpu --no-parity --no-e2e > /dev/null 2>> gpurun_out/${T}_ncu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_csr_fused -s 12 -c 1 -o gpurun_out/${T}_k1_fused -f \
  python bench.py --steps 5 --warmup 1 --also "" --no-cpu --no-parity --no-e2e > /dev/null 2>> gpurun_out/${T}_ncu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_dense -s 20 -c 1 -o gpurun_out/${T}_k1_dense -f \
  python bench.py --workload cfg2 --steps 5 --warmup 1 --also "" --no-cpu --no-parity --no-e2e > /dev/null 2>> gpurun_out/${T}_ncu.err
ls -la gpurun_out | tail -12
