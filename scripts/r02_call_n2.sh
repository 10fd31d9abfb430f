#!/bin/bash
# round 2, 2-GPU call: N-rank gradients == full-batch gradients (NCCL), weak scaling N=1 vs N=2, overlap A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/ddp_grad_check.py > gpurun_out/r02n2_ddp_check.log 2>&1; echo "[ddp grad check exit $?]"; grep -E "DDP_GRAD_CHECK|Error|error|assert" gpurun_out/r02n2_ddp_check.log | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02n2_bench_n1.json 2> gpurun_out/r02n2_bench_n1.err; echo "[bench n1 exit $?]"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --overlap 1 > gpurun_out/r02n2_bench_n2.json 2> gpurun_out/r02n2_bench_n2.err; echo "[bench n2 exit $?]"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --overlap 2 > gpurun_out/r02n2_bench_n2_overlap.json 2> gpurun_out/r02n2_bench_n2_overlap.err; echo "[bench n2 overlap exit $?]"
python - <<'PY'
import json
def val(f):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); return d["value"], d["ms_per_step"], d["e2e"]["value"]
    except Exception as e:
        return (None, None, repr(e)[:100])
n1 = val("gpurun_out/r02n2_bench_n1.json"); n2 = val("gpurun_out/r02n2_bench_n2.json"); n2o = val("gpurun_out/r02n2_bench_n2_overlap.json")
print("N=1", n1); print("N=2 one all-reduce", n2); print("N=2 two all-reduces (1 overlapped)", n2o)
if n1[0] and n2[0]: print("efficiency N=2:", n2[0] / (2 * n1[0]), " overlapped:", (n2o[0] or 0) / (2 * n1[0]))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 --overlap 3 > gpurun_out/r02n2_bench_n2_overlap3.json 2> gpurun_out/r02n2_bench_n2_overlap3.err; echo "[bench n2 overlap3 exit $?]"
python scripts/summarize.py gpurun_out/r02n2_bench_n2_overlap3.json | grep -E "value|ms_per" | head -3
tail -3 gpurun_out/r02n2_bench_n2.err
