#!/bin/bash
# round 2, session 2: phase trace of attn_bwd_rows_kernel (first-wave CTA and a mid-grid CTA) at the c3-max and config-2 shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CREAM_B200_LIB=build_trace/libcream_b200_trace.so CREAM_ATTN_TRACE=1 timeout 300 python scripts/time_attention.py > gpurun_out/r02n_trace.log 2>&1
grep -A6 "ROWS TRACE" gpurun_out/r02n_trace.log | head -80
grep -E "fwd" gpurun_out/r02n_trace.log
