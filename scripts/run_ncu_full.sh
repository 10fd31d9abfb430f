#!/bin/bash
# one `ncu --set full` capture per kernel family (single GPU; short command)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "ln_bwd:40" "attn_bwd_rows_kernel:14" "attn_bwd_cols_kernel:14" "attn_fwd_kernel:30" "gemm_bf16_kernel:330"; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c ${CNT:-2} -o gpurun_out/prof_$k -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "[ncu $k exit $?]"
done
ls -la gpurun_out | head -30
