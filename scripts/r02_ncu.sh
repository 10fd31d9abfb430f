#!/bin/bash
# round 2: ncu launch list of the bench command + one `--set full` capture per kernel class (single GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1; echo "[ncu launch list exit $?]"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 60 -c 6 -o gpurun_out/prof_r02_attention -f $B > gpurun_out/r02_ncu_attn.log 2>&1; echo "[ncu attention exit $?]"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ln_|colsum|cast_scale|adamw_kernel|im2col|assemble|xent" -s 120 -c 12 -o gpurun_out/prof_r02_bytemovers -f $B > gpurun_out/r02_ncu_bytes.log 2>&1; echo "[ncu byte movers exit $?]"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 420 -c 14 -o gpurun_out/prof_r02_gemm -f $B > gpurun_out/r02_ncu_gemm.log 2>&1; echo "[ncu gemm exit $?]"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_|rpe_index" -s 4 -c 10 -o gpurun_out/prof_r02_c2 -f python scripts/profile_c2_kernels.py > gpurun_out/r02_ncu_c2.log 2>&1; echo "[ncu c2 kernels exit $?]"
# summaries are produced ON THE BOX (ncu -i needs no GPU); the reports themselves exceed the 64 MiB return limit
python scripts/r02_make_profiles.py gpurun_out/r02_profiles > gpurun_out/r02_make_profiles.log 2>&1; echo "[summaries exit $?]"
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/prof_r02_bytemovers.ncu-rep gpurun_out/prof_r02_c2.ncu-rep gpurun_out/prof_r02_gemm.ncu-rep
ls gpurun_out/r02_profiles
