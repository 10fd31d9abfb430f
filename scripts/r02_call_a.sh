#!/bin/bash
# round 2, GPU call A: the whole GPU suite (old + BASELINE-config parity + reference-unchanged), smoke, bench c3 / c5.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r02a_gpu_tests.log 2>&1; echo "[pytest exit $?]"
grep -E "passed|failed|FAILED|ERROR|^\[|Error" gpurun_out/r02a_gpu_tests.log | tail -70
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "[smoke exit $?]"; tail -2 gpurun_out/r02a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_c3.json 2> gpurun_out/r02a_bench_c3.err; echo "[bench c3 exit $?]"; tail -c 2500 gpurun_out/r02a_bench_c3.json; tail -3 gpurun_out/r02a_bench_c3.err
timeout 900 python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_c5.json 2> gpurun_out/r02a_bench_c5.err; echo "[bench c5 exit $?]"; tail -c 1500 gpurun_out/r02a_bench_c5.json; tail -3 gpurun_out/r02a_bench_c5.err
