#!/bin/bash
# round 2: TinyCLIP towers - GPU parity against the reference model, then the config-4 bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_clip.py -m gpu -q -s -x -p no:cacheprovider > gpurun_out/r02g_gpu_clip.log 2>&1; echo "[pytest clip exit $?]"
grep -E "passed|failed|FAILED|ERROR|Error|worst|assert" gpurun_out/r02g_gpu_clip.log | tail -25
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/r02g_bench_c4.json 2> gpurun_out/r02g_bench_c4.err; echo "[bench c4 exit $?]"
python scripts/summarize.py gpurun_out/r02g_bench_c4.json | cut -c1-1200; tail -5 gpurun_out/r02g_bench_c4.err
du -sh gpurun_out
timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench_c2.json 2> gpurun_out/r02g_bench_c2.err; echo "[bench c2 exit $?]"
python scripts/summarize.py gpurun_out/r02g_bench_c2.json | grep -E "value|ms_per|roofline|speedup" | cut -c1-700
