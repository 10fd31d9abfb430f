#!/bin/bash
# round 2, session 2: 256-bit row stores + cheaper bucket-row staging in the attention backward: parity, timing, phase trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_native.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python scripts/time_attention.py 2>&1 | tee gpurun_out/r02p_time_attention.log | grep -E "fwd"
CREAM_B200_LIB=build_trace/libcream_b200_trace.so CREAM_ATTN_TRACE=1 CREAM_ONLY_STRUCTURED=1 timeout 300 python scripts/time_attention.py > gpurun_out/r02p_trace.log 2>&1
grep -A6 "ROWS TRACE cta mid" gpurun_out/r02p_trace.log | head -14
for v in 1 1; do
  CREAM_PDL=$v timeout 600 python bench.py --quick --steps 20 --warmup 5 2>gpurun_out/r02p_bench.err | tee -a gpurun_out/r02p_bench.jsonl | cut -c1-250
done
