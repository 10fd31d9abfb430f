#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_native.py tests/test_gpu_clip.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
