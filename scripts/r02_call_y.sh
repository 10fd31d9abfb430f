#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== PDL off"; CREAM_PDL=0 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_native.py tests/test_gpu_clip.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
echo "== PDL on, shadows cleared before every test"; CREAM_TEST_CLEAR_SHADOWS=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_native.py tests/test_gpu_clip.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
