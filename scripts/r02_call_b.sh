#!/bin/bash
# round 2, GPU call B: native-runtime tests (incl. grid-product attention path), the failing cases of call A,
# bench c3 / c2, launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_native.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02b_native.log 2>&1; echo "[pytest native exit $?]"
grep -E "passed|failed|FAILED|ERROR|^\[|Error|error|assert" gpurun_out/r02b_native.log | tail -40
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -s -p no:cacheprovider -k "config2 or trainer or large_magnitude" > gpurun_out/r02b_baseline.log 2>&1; echo "[pytest baseline exit $?]"
grep -E "passed|failed|FAILED|ERROR|^\[|Error|error|assert" gpurun_out/r02b_baseline.log | tail -40
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "irpe or attention" > gpurun_out/r02b_parity_attn.log 2>&1; echo "[pytest parity-attention exit $?]"
tail -5 gpurun_out/r02b_parity_attn.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench_c3.err; echo "[bench c3 exit $?]"; python scripts/summarize.py gpurun_out/r02b_bench_c3.json 2>/dev/null || tail -c 1800 gpurun_out/r02b_bench_c3.json; tail -3 gpurun_out/r02b_bench_c3.err
timeout 900 python bench.py --python-engine --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_c3_py.json 2> gpurun_out/r02b_bench_c3_py.err; echo "[bench c3 python-engine exit $?]"; python scripts/summarize.py gpurun_out/r02b_bench_c3_py.json 2>/dev/null; tail -3 gpurun_out/r02b_bench_c3_py.err
timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err; echo "[bench c2 exit $?]"; tail -c 2500 gpurun_out/r02b_bench_c2.json; tail -3 gpurun_out/r02b_bench_c2.err
timeout 600 python scripts/time_attention.py > gpurun_out/r02b_time_attention.log 2>&1; tail -12 gpurun_out/r02b_time_attention.log
