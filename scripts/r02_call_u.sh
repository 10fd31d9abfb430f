#!/bin/bash
# round 2, session 2, 2-GPU call: TinyViT fused bias gather + window packing (1 GPU), then N = 1 vs N = 2 with programmatic
# dependent launch on (weak scaling, one all-reduce after the backward) and the N-rank == full-batch gradient check.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "tinyvit or packed_items or irpe or rpe_attention" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/ddp_grad_check.py > gpurun_out/r02u_ddp_check.log 2>&1; echo "[ddp grad check exit $?]"; grep -E "DDP_GRAD_CHECK|Error|error|assert" gpurun_out/r02u_ddp_check.log | tail -5
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --quick --steps 20 --warmup 5 2>gpurun_out/r02u_bench_n1.err | tee gpurun_out/r02u_bench_n1.json | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --quick --steps 20 --warmup 5 2>gpurun_out/r02u_bench_n2.err | tee gpurun_out/r02u_bench_n2.json | cut -c1-200
python - <<'PY'
import json
def val(f):
    try:
        return json.loads(open(f).read().strip().splitlines()[-1])["value"]
    except Exception as e:
        return None
a, b = val("gpurun_out/r02u_bench_n1.json"), val("gpurun_out/r02u_bench_n2.json")
print("N=1", a, "N=2", b, "efficiency", (b / (2 * a)) if a and b else None)
PY
tail -3 gpurun_out/r02u_bench_n2.err
