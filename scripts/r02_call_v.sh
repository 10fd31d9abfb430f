#!/bin/bash
# round 2, session 2: packed-fp32 (FFMA2 / FADD2) column loop of the AutoFormer attention backward; TinyViT fused bias gather
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_native.py tests/test_gpu_baseline_configs.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for v in 1 0; do
  echo "== CREAM_AF_F32X2=$v"
  CREAM_AF_F32X2=$v CREAM_ONLY_STRUCTURED=1 timeout 300 python scripts/time_attention.py 2>&1 | grep -E "structured"
  CREAM_AF_F32X2=$v timeout 600 python bench.py --quick --steps 20 --warmup 5 2>gpurun_out/r02v_bench.err | tee -a gpurun_out/r02v_bench.jsonl | cut -c1-200
done
CREAM_B200_LIB=build_trace/libcream_b200_trace.so CREAM_ATTN_TRACE=1 CREAM_ONLY_STRUCTURED=1 timeout 300 python scripts/time_attention.py > gpurun_out/r02v_trace.log 2>&1
grep -A6 "ROWS TRACE cta mid" gpurun_out/r02v_trace.log | head -7
