#!/bin/bash
# round 2, final validation of the session: whole GPU suite, smoke, bench lines (c3 headline with every leg, c5, c2, reference arm),
# ncu launch list + `--set full` captures of the attention and GEMM kernels inside the bench step, summaries made on the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_gpu_suite.log 2>&1; echo "[pytest -m gpu exit $?]"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r02_gpu_suite.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "[smoke exit $?]"; tail -1 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo "[bench c3 exit $?]"
python scripts/summarize.py gpurun_out/r02_bench_c3.json | grep -E "^ *(value|ms_per_step|e2e|speedup|gpu_time_share|roofline)" | cut -c1-500
timeout 600 python bench.py --config c5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; echo "[bench c5 exit $?]"
python scripts/summarize.py gpurun_out/r02_bench_c5.json | grep -E "^ *(value|ms_per_step|e2e|speedup)" | cut -c1-300
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; echo "[bench c2 exit $?]"
python scripts/summarize.py gpurun_out/r02_bench_c2.json | grep -E "^ *(value|ms_per_step|e2e|speedup|roofline)" | cut -c1-500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "[reference arm exit $?]"
timeout 300 python scripts/time_attention.py > gpurun_out/r02_time_attention.log 2>&1; grep fwd gpurun_out/r02_time_attention.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1; echo "[ncu launch list exit $?]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 60 -c 6 -o gpurun_out/prof_r02_attention -f $B > gpurun_out/r02_ncu_attn.log 2>&1; echo "[ncu attention exit $?]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 420 -c 14 -o gpurun_out/prof_r02_gemm -f $B > gpurun_out/r02_ncu_gemm.log 2>&1; echo "[ncu gemm exit $?]"
python scripts/r02_make_profiles.py gpurun_out/r02_profiles > gpurun_out/r02_make_profiles.log 2>&1; echo "[summaries exit $?]"
rm -f gpurun_out/prof_r02_gemm.ncu-rep gpurun_out/prof_r02_attention.ncu-rep
ls gpurun_out/r02_profiles
# timing ablations of the rows kernel's column loop (debug builds; results wrong by construction): 1 = no workspace stores,
# 9 = + no MUFU, 25 = + no packed-dT TMEM store, 29 = + no bucket accumulations
for a in abl1 abl9 abl25 abl29; do
  echo "== $a"; CREAM_B200_LIB=build_trace/libcream_b200_$a.so CREAM_ATTN_TRACE=1 CREAM_ONLY_STRUCTURED=1 timeout 200 python scripts/time_attention.py 2>&1 | grep -E "structured|mid-grid" -A4 | grep -E "structured|row thread half 0" | tail -2
done
