#!/bin/bash
# bench line + ncu launch list + one full capture of the top kernels (GPU box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "[smoke exit $?]"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "[bench exit $?]"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 900 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "[ncu launches exit $?]"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 30 -c 2 -o gpurun_out/prof_attn_fwd -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1; echo "[ncu attn exit $?]"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 300 -c 4 -o gpurun_out/prof_gemm -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; echo "[ncu gemm exit $?]"
fi
ls -la gpurun_out
