#!/bin/bash
# round 2, GPU call C: tensor-core structured attention (CREAM_AF_MMA) + grid-product fix + flat AdamW v2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/debug_gridprod.py > gpurun_out/r02c_debug_gridprod.log 2>&1; tail -12 gpurun_out/r02c_debug_gridprod.log
for mode in 1 0; do
  CREAM_AF_MMA=$mode timeout 600 python scripts/time_attention.py > gpurun_out/r02c_time_attention_mma$mode.log 2>&1; echo "[time_attention AF_MMA=$mode]"; tail -5 gpurun_out/r02c_time_attention_mma$mode.log
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "attention or supernet or irpe" > gpurun_out/r02c_parity.log 2>&1; echo "[pytest parity exit $?]"; tail -8 gpurun_out/r02c_parity.log
timeout 900 python -m pytest tests/test_gpu_native.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02c_native.log 2>&1; echo "[pytest native exit $?]"
grep -E "passed|failed|FAILED|ERROR|^\[|Error|assert" gpurun_out/r02c_native.log | tail -25
timeout 1200 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02c_baseline.log 2>&1; echo "[pytest baseline exit $?]"
grep -E "passed|failed|FAILED|ERROR|^\[|Error|assert" gpurun_out/r02c_baseline.log | tail -40
for mode in 1 0; do
  CREAM_AF_MMA=$mode timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_c3_mma$mode.json 2> gpurun_out/r02c_bench_c3_mma$mode.err; echo "[bench c3 AF_MMA=$mode exit $?]"
  python scripts/summarize.py gpurun_out/r02c_bench_c3_mma$mode.json | cut -c1-1500; tail -3 gpurun_out/r02c_bench_c3_mma$mode.err
done
timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_c2.json 2> gpurun_out/r02c_bench_c2.err; echo "[bench c2 exit $?]"; python scripts/summarize.py gpurun_out/r02c_bench_c2.json | cut -c1-1200; tail -3 gpurun_out/r02c_bench_c2.err
