#!/bin/bash
# bench at N=1 and N=NG (torchrun, one rank per GPU over NCCL) on one box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=${NG:-2}
nvidia-smi -L
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "[bench n1 exit $?]"; tail -c 1200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "[bench n$NG exit $?]"; tail -c 1500 gpurun_out/bench_n$NG.json; tail -5 gpurun_out/bench_n$NG.err
