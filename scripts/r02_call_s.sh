#!/bin/bash
# round 2, session 2: the TinyCLIP step as one CUDA graph
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clip.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_bench_c4.json 2> gpurun_out/r02s_bench_c4.err; echo "[bench c4 graph exit $?]"
tail -3 gpurun_out/r02s_bench_c4.err
python scripts/summarize.py gpurun_out/r02s_bench_c4.json | grep -E "value|ms_per|e2e|speedup|engine|host_enqueue|gpu_launches" | cut -c1-300
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r02s_bench_c4_eager.json 2> gpurun_out/r02s_bench_c4_eager.err; echo "[bench c4 eager exit $?]"
python scripts/summarize.py gpurun_out/r02s_bench_c4_eager.json | grep -E "value|ms_per|e2e|speedup|host_enqueue" | cut -c1-300
