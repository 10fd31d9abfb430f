#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "layernorm" -p no:cacheprovider 2>&1 | tail -3
timeout 600 python scripts/tune_ln_bwd.py 2>&1 | tee gpurun_out/r02l_tune_ln_bwd.log | tail -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_pipe -s 20 -c 2 -o gpurun_out/prof_r02_lnpipe -f python scripts/tune_ln_bwd.py > gpurun_out/r02l_ncu.log 2>&1; echo "[ncu exit $?]"
mkdir -p gpurun_out/r02_lnpipe; python scripts/r02_make_profiles.py gpurun_out/r02_lnpipe > /dev/null 2>&1; rm -f gpurun_out/prof_r02_lnpipe.ncu-rep
cat gpurun_out/r02_lnpipe/r02_ncu_lnpipe.txt | cut -c1-330
