#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "layernorm" -p no:cacheprovider 2>&1 | tail -5
timeout 600 python scripts/time_bytemovers.py 2>&1 | tee gpurun_out/r02k_time_bytemovers.log | tail -120
timeout 900 python -m pytest tests/test_gpu_native.py tests/test_gpu_baseline_configs.py tests/test_gpu_clip.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02k_bench_c3.json 2> gpurun_out/r02k_bench_c3.err; echo "[bench c3 exit $?]"
python scripts/summarize.py gpurun_out/r02k_bench_c3.json | grep -E "value|ms_per|e2e|byte_movers|gpu_time_share" | cut -c1-900
