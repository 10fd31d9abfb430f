#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/debug_deit.py > gpurun_out/r02d_debug_deit.log 2>&1; grep -E "<<<<|per-bucket|ratio" gpurun_out/r02d_debug_deit.log | cut -c1-700
for mode in 0 1; do
  CREAM_AF_MMA=$mode timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 3 -c 3 -o gpurun_out/prof_r02d_af_mma$mode -f python scripts/profile_af_kernels.py > gpurun_out/r02d_ncu_af_mma$mode.log 2>&1; echo "[ncu AF_MMA=$mode exit $?]"
done
timeout 900 python -m pytest tests/test_gpu_native.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02d_native.log 2>&1; echo "[pytest native exit $?]"
grep -E "passed|failed|FAILED|ERROR|^\[|Error|assert" gpurun_out/r02d_native.log | tail -25
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench_c3.json 2> gpurun_out/r02d_bench_c3.err; echo "[bench c3 exit $?]"
python scripts/summarize.py gpurun_out/r02d_bench_c3.json | grep -v "kernel_table\|roofline\|attention\|byte_movers" | cut -c1-600; tail -2 gpurun_out/r02d_bench_c3.err
