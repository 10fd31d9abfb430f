#!/bin/bash
# Runs every native C-ABI test case in its own process with a timeout (GPU box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/native.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
CASES=${@:-$(./build/test_native list)}
for c in $CASES; do
  timeout 120 ./build/test_native $c >> $LOG 2>&1
  echo "[exit $?] $c" >> $LOG
done
grep -E "PASS|FAIL|exit|TFLOP|GB/s|timeout|error" $LOG | tail -80
