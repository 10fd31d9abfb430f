"""LayerNorm forward / backward, bf16 cast and bias-gradient column sums timed alone (CUDA events, rotating buffer sets
larger than the L2) at the supernet-S shapes: achieved GB/s of algorithmic bytes against the measured copy peak."""
import json
import os
import sys
from pathlib import Path
import torch
sys.path.insert(0, ".")
from cream_b200 import ops

dev = "cuda"
peaks = json.loads(Path("MEASURED_PEAKS.json").read_text()) if Path("MEASURED_PEAKS.json").exists() else {}
hbm = peaks.get("hbm_gbs", 6572.0)
M = 128 * 197
SETS = 4


def timeit(fns, reps=5):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(6_000_000)        # ~3 ms: the launches below queue up behind it, so the host is not in the timing
    e0.record()
    for _ in range(reps):
        for f in fns:
            f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns)) * 1e-3


for E in (320, 384, 448, 624):
    xs = [torch.randn(M, E, device=dev) for _ in range(SETS)]
    rg = [torch.randn(M, E, device=dev) for _ in range(SETS)]
    gam, bet = torch.randn(E, device=dev), torch.randn(E, device=dev)
    st = [ops.layernorm_fwd(x, gam, bet, 1e-5, E) for x in xs]
    dys = [s[0] for s in st]
    dg, db = torch.zeros(E, device=dev), torch.zeros(E, device=dev)
    t = timeit([lambda i=i: ops.layernorm_fwd(xs[i], gam, bet, 1e-5, E) for i in range(SETS)])
    by = M * E * 6.0
    print(f"E {E:4d}  ln_fwd      {t * 1e6:7.1f} us  {by / t / 1e9:7.0f} GB/s  {by / t / 1e9 / hbm:5.2f} of peak")
    t = timeit([lambda i=i: ops.layernorm_bwd(dys[i], xs[i], gam, st[i][1], st[i][2], E, dg, db, resid_grad=rg[i]) for i in range(SETS)])
    by = M * E * 14.0
    print(f"E {E:4d}  ln_bwd      {t * 1e6:7.1f} us  {by / t / 1e9:7.0f} GB/s  {by / t / 1e9 / hbm:5.2f} of peak")
    dbias = torch.zeros(E, device=dev)
    by = M * E * 6.0
    for u in (2, 4, 8):
        for bps in (0, 4, 8):
            os.environ["CREAM_TUNE_CAST_U"], os.environ["CREAM_TUNE_CAST_BPS"] = str(u), str(bps)
            t = timeit([lambda i=i: ops.cast_scale(rg[i], dbias=dbias) for i in range(SETS)])
            print(f"E {E:4d}  cast_scale U{u} BPS{bps}  {t * 1e6:7.1f} us  {by / t / 1e9:7.0f} GB/s  {by / t / 1e9 / hbm:5.2f} of peak")
    os.environ.pop("CREAM_TUNE_CAST_U"); os.environ.pop("CREAM_TUNE_CAST_BPS")
    del xs, rg, st, dys
for F in (1344, 1792):
    hs = [ops.empty_bf16(M, F, dev) for _ in range(SETS)]
    for h in hs:
        h.copy_(torch.randn(M, F, device=dev))
    dbf = torch.zeros(F, device=dev)
    by = M * F * 2.0
    for u in (2, 4):
        for bps in (0, 4, 8):
            os.environ["CREAM_TUNE_COLSUM_U"], os.environ["CREAM_TUNE_COLSUM_BPS"] = str(u), str(bps)
            t = timeit([lambda i=i: ops.bias_grad(hs[i], dbf) for i in range(SETS)])
            print(f"F {F:4d}  bias_grad U{u} BPS{bps}  {t * 1e6:7.1f} us  {by / t / 1e9:7.0f} GB/s  {by / t / 1e9 / hbm:5.2f} of peak")
    os.environ.pop("CREAM_TUNE_COLSUM_U"); os.environ.pop("CREAM_TUNE_COLSUM_BPS")
    del hs
