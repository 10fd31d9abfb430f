#!/bin/bash
# round 2: one `ncu --set full` launch of EVERY kernel class at BASELINE shapes (scripts/profile_kernel_classes.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02_classes
rm -f gpurun_out/prof_r02_*.ncu-rep
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r02_classes -f \
    python scripts/profile_kernel_classes.py > gpurun_out/r02_ncu_classes.log 2>&1; echo "[ncu classes exit $?]"
grep -A60 "^ORDER" gpurun_out/r02_ncu_classes.log > gpurun_out/r02_classes/order.txt
python scripts/r02_make_profiles.py gpurun_out/r02_classes > gpurun_out/r02_make_classes.log 2>&1; echo "[summary exit $?]"
ls -la gpurun_out/*.ncu-rep; rm -f gpurun_out/prof_r02_classes.ncu-rep
ls gpurun_out/r02_classes; head -50 gpurun_out/r02_classes/r02_ncu_classes.txt | cut -c1-260
