#!/bin/bash
# round 2, GPU call A: the whole GPU suite (old + BASELINE-config parity + reference-unchanged + native runtime),
# smoke, bench c3 / c5 / c2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt
for f in test_gpu_native test_gpu_baseline_configs test_gpu_reference_unchanged test_gpu_parity; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02a_$f.log 2>&1; echo "[pytest $f exit $?]"
  grep -E "passed|failed|FAILED|ERROR|^\[|Error|error" gpurun_out/r02a_$f.log | tail -40
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "[smoke exit $?]"; tail -2 gpurun_out/r02a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_c3.json 2> gpurun_out/r02a_bench_c3.err; echo "[bench c3 exit $?]"; tail -c 2500 gpurun_out/r02a_bench_c3.json; tail -3 gpurun_out/r02a_bench_c3.err
timeout 900 python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_c5.json 2> gpurun_out/r02a_bench_c5.err; echo "[bench c5 exit $?]"; tail -c 1500 gpurun_out/r02a_bench_c5.json; tail -3 gpurun_out/r02a_bench_c5.err
timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err; echo "[bench c2 exit $?]"; tail -c 2500 gpurun_out/r02a_bench_c2.json; tail -3 gpurun_out/r02a_bench_c2.err
