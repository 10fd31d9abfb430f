#!/bin/bash
# GPU parity tests, one process per group so a device trap cannot poison later groups.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/gpu_tests.log
: > $LOG
GROUPS_=${@:-"rpe_index sliced_or_gemm layernorm attention_autoformer attention_full_or_structured irpe supernet_vs_golden fused_matches supernet_s cpu_tensors dense_logit clip_attention tinyvit"}
for g in $GROUPS_; do
  k=$(echo $g | sed 's/_or_/ or /g')
  echo "===== group: $k" >> $LOG
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$k" -p no:cacheprovider -s >> $LOG 2>&1
  echo "[exit $?] $g" >> $LOG
done
grep -E "^=====|passed|failed|\[exit|Error|error|rel err|assert " $LOG | tail -120
