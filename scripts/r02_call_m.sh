#!/bin/bash
# round 2: validation after the LayerNorm-backward pipeline (+ fused bf16 emit), side-stream bias gradients and the
# TMA-staged delta in the attention backward: whole GPU suite, attention timings, bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02m_gpu_suite.log 2>&1; echo "[pytest -m gpu exit $?]"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r02m_gpu_suite.log | tail -12
timeout 600 python scripts/time_attention.py > gpurun_out/r02m_time_attention.log 2>&1; tail -5 gpurun_out/r02m_time_attention.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench_c3.json 2> gpurun_out/r02m_bench_c3.err; echo "[bench c3 exit $?]"
python scripts/summarize.py gpurun_out/r02m_bench_c3.json | grep -E "value|ms_per|e2e|speedup|gpu_time_share" | cut -c1-600
timeout 900 python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02m_bench_c5.json 2> gpurun_out/r02m_bench_c5.err; echo "[bench c5 exit $?]"
python scripts/summarize.py gpurun_out/r02m_bench_c5.json | grep -E "value|ms_per|e2e|speedup" | cut -c1-300
timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench_c2.json 2> gpurun_out/r02m_bench_c2.err; echo "[bench c2 exit $?]"
python scripts/summarize.py gpurun_out/r02m_bench_c2.json | grep -E "value|ms_per|e2e|speedup" | cut -c1-300
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench_c4.json 2> gpurun_out/r02m_bench_c4.err; echo "[bench c4 exit $?]"
python scripts/summarize.py gpurun_out/r02m_bench_c4.json | grep -E "value|ms_per|e2e|speedup" | cut -c1-300
