#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_native.py tests/test_gpu_clip.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02x_order.log 2>&1; echo "[exit $?]"
grep -n "Error\|assert\|^E " gpurun_out/r02x_order.log | head -30
tail -5 gpurun_out/r02x_order.log
timeout 300 python -m pytest tests/test_gpu_clip.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
