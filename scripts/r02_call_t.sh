#!/bin/bash
# round 2, session 2: two 50-token items per attention tile (block-diagonal visibility) for the TinyCLIP image tower
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "packed_items or attention" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_clip.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_bench_c4.json 2> gpurun_out/r02t_bench_c4.err; echo "[bench c4 exit $?]"
python scripts/summarize.py gpurun_out/r02t_bench_c4.json | grep -E "value|ms_per|e2e|speedup|per_kind" | cut -c1-600
