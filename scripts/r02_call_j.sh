#!/bin/bash
# round 2, 2-GPU call for TinyCLIP (config 4): N-rank gradients == full-batch gradients over NCCL (feature all_gather with
# gradient), then weak scaling N=1 vs N=2 of the contrastive step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/clip_ddp_check.py > gpurun_out/r02j_clip_ddp_check.log 2>&1; echo "[clip ddp check exit $?]"; grep -E "CLIP_DDP_CHECK|Error|error|assert" gpurun_out/r02j_clip_ddp_check.log | tail -8
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_c4_n1.json 2> gpurun_out/r02j_bench_c4_n1.err; echo "[bench c4 n1 exit $?]"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --config c4 --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02j_bench_c4_n2.json 2> gpurun_out/r02j_bench_c4_n2.err; echo "[bench c4 n2 exit $?]"
python scripts/summarize.py gpurun_out/r02j_bench_c4_n1.json gpurun_out/r02j_bench_c4_n2.json | grep -E "==|value|ms_per|e2e|speedup" | cut -c1-400
tail -3 gpurun_out/r02j_bench_c4_n2.err
