#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 1; do CREAM_PDL=$v timeout 300 python scripts/repeat_staged_test.py 6 2>&1 | grep -v Warning | tee -a gpurun_out/r02q_staged.log; done
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02q_gpu_suite.log 2>&1; echo "[pytest -m gpu exit $?]"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r02q_gpu_suite.log | tail -12
for v in 0 1; do
  CREAM_SIDE_WGRAD=$v timeout 600 python bench.py --quick --steps 20 --warmup 5 2>gpurun_out/r02q_bench.err | tee -a gpurun_out/r02q_bench.jsonl | cut -c1-200
done
