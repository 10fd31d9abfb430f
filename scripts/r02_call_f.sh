#!/bin/bash
# round 2, final GPU call: ncu captures + summaries, then the whole GPU suite, smoke and the bench lines.
cd "$(dirname "$0")/.."
bash scripts/r02_ncu.sh
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02f_gpu_suite.log 2>&1; echo "[pytest -m gpu exit $?]"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r02f_gpu_suite.log | tail -12
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02f_gpu_parity_comparator.log 2>&1; echo "[comparator log exit $?]"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "[smoke exit $?]"; tail -1 gpurun_out/r02f_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_c3.json 2> gpurun_out/r02f_bench_c3.err; echo "[bench c3 exit $?]"
python scripts/summarize.py gpurun_out/r02f_bench_c3.json | grep -v "kernel_table" | cut -c1-900; tail -2 gpurun_out/r02f_bench_c3.err
timeout 900 python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_c5.json 2> gpurun_out/r02f_bench_c5.err; echo "[bench c5 exit $?]"
python scripts/summarize.py gpurun_out/r02f_bench_c5.json | grep -E "value|ms_per|e2e|tflops" | cut -c1-300
timeout 900 python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/r02f_bench_c2.json 2> gpurun_out/r02f_bench_c2.err; echo "[bench c2 exit $?]"
python scripts/summarize.py gpurun_out/r02f_bench_c2.json | grep -E "value|ms_per|e2e|gpu_reference|speedup|cpu_baseline|roofline" | cut -c1-700
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2> gpurun_out/r02f_bench_reference_arm.err; echo "[reference arm exit $?]"
timeout 600 python scripts/time_attention.py > gpurun_out/r02f_time_attention.log 2>&1; tail -5 gpurun_out/r02f_time_attention.log
du -sh gpurun_out
