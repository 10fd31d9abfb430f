#!/bin/bash
# round 2, session 2: (1) timing ablations of attn_bwd_rows_kernel's store paths, (2) programmatic dependent launch A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for a in abl1 abl3 abl7; do
  echo "== $a"; CREAM_B200_LIB=build_trace/libcream_b200_$a.so CREAM_ONLY_STRUCTURED=1 timeout 200 python scripts/time_attention.py 2>&1 | grep -E "structured|Error|error" | head -3
done
echo "== PDL parity subset"
timeout 900 python -m pytest tests/test_gpu_native.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for v in 0 1 0 1; do
  CREAM_PDL=$v timeout 600 python bench.py --quick --steps 20 --warmup 5 2>gpurun_out/r02o_bench_pdl$v.err | tee -a gpurun_out/r02o_bench_pdl.jsonl | cut -c1-400
done
