#!/bin/bash
# Runs every native C-ABI test case in its own process with a timeout (GPU box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/native.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
CASES=${@:-$(./build/test_native list)}
for c in $CASES; do
  case $c in
    gemm_*|perf_*)   # every GEMM case twice: one CTA per 128-row tile, and CTA pairs (cta_group::2)
      for mode in 1 2; do
        echo "-- cta_pair=$mode" >> $LOG
        CREAM_TEST_CTA_PAIR=$mode timeout 120 ./build/test_native $c >> $LOG 2>&1
        echo "[exit $?] $c cta_pair=$mode" >> $LOG
      done ;;
    *)
      timeout 120 ./build/test_native $c >> $LOG 2>&1
      echo "[exit $?] $c" >> $LOG ;;
  esac
done
grep -E "PASS|FAIL|exit|TFLOP|GB/s|timeout|error" $LOG | tail -80
