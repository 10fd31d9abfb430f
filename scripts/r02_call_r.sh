#!/bin/bash
# round 2, session 2: staged TMA stores of the attention-backward workspace (CREAM_BWD_STAGE) + split dQ MMA
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 1 0; do
  echo "== CREAM_BWD_STAGE=$v"
  CREAM_BWD_STAGE=$v timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_native.py -m gpu -q -x -p no:cacheprovider -k "attention or grid_product or deit or irpe or rpe" 2>&1 | tail -3
  CREAM_BWD_STAGE=$v CREAM_ONLY_STRUCTURED= timeout 300 python scripts/time_attention.py 2>&1 | grep -E "structured|c2 struct"
done
CREAM_B200_LIB=build_trace/libcream_b200_trace.so CREAM_ATTN_TRACE=1 CREAM_ONLY_STRUCTURED=1 timeout 300 python scripts/time_attention.py > gpurun_out/r02r_trace.log 2>&1
grep -A6 "ROWS TRACE cta mid" gpurun_out/r02r_trace.log | head -7
for v in 1 0; do
  CREAM_BWD_STAGE=$v timeout 600 python bench.py --quick --steps 20 --warmup 5 2>gpurun_out/r02r_bench.err | tee -a gpurun_out/r02r_bench.jsonl | cut -c1-200
done
