#!/bin/bash
# round 2: headline bench lines with the GPU-side reference leg (reference model, stock PyTorch ops, same GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench_c3.json 2> gpurun_out/r02h_bench_c3.err; echo "[bench c3 exit $?]"
python scripts/summarize.py gpurun_out/r02h_bench_c3.json | grep -E "value|ms_per|e2e|gpu_reference|speedup|roofline" | cut -c1-500; tail -3 gpurun_out/r02h_bench_c3.err
timeout 900 python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_c5.json 2> gpurun_out/r02h_bench_c5.err; echo "[bench c5 exit $?]"
python scripts/summarize.py gpurun_out/r02h_bench_c5.json | grep -E "value|ms_per|e2e|gpu_reference|speedup" | cut -c1-500; tail -3 gpurun_out/r02h_bench_c5.err
