#!/bin/bash
# round 2, session 2: last sanity pass over the committed tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_native.py tests/test_gpu_clip.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
CREAM_ONLY_STRUCTURED=1 timeout 300 python scripts/time_attention.py 2>&1 | grep -E "structured"
timeout 600 python bench.py --quick --steps 20 --warmup 5 2>gpurun_out/r02w_bench.err | tee gpurun_out/r02w_bench.jsonl | cut -c1-220
timeout 600 python bench.py --quick --config c5 --steps 20 --warmup 5 2>>gpurun_out/r02w_bench.err | tee -a gpurun_out/r02w_bench.jsonl | cut -c1-220
