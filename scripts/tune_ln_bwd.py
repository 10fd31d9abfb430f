"""LayerNorm backward variants timed on the GPU clock (launches queued behind a spin kernel): register-resident
kernel vs the bulk-copy pipeline, per-row vs per-tile copies, ring depth; with and without the fused bf16 emit."""
import os
import sys
import torch
sys.path.insert(0, ".")
from cream_b200 import ops

dev, M, SETS = "cuda", 128 * 197, 4


def timeit(fns, reps=5):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(6_000_000)
    e0.record()
    for _ in range(reps):
        for f in fns:
            f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns)) * 1e3


for E in (320, 448, 624):
    xs = [torch.randn(M, E, device=dev) for _ in range(SETS)]
    rg = [torch.randn(M, E, device=dev) for _ in range(SETS)]
    gam, bet = torch.randn(E, device=dev), torch.randn(E, device=dev)
    st = [ops.layernorm_fwd(x, gam, bet, 1e-5, E) for x in xs]
    dg, db, dbias = torch.zeros(E, device=dev), torch.zeros(E, device=dev), torch.zeros(E, device=dev)
    plain = [lambda i=i: ops.layernorm_bwd(st[i][0], xs[i], gam, st[i][1], st[i][2], E, dg, db, resid_grad=rg[i]) for i in range(SETS)]
    fused = [lambda i=i: ops.layernorm_bwd_cast(st[i][0], xs[i], gam, st[i][1], st[i][2], E, dg, db, resid_grad=rg[i], dbias=dbias) for i in range(SETS)]
    cast = [lambda i=i: ops.cast_scale(rg[i], dbias=dbias) for i in range(SETS)]
    for name, env in (("register-resident kernel", dict(CREAM_LN_PIPE="0")), ("bulk-copy pipeline", dict())):
        os.environ.pop("CREAM_LN_PIPE", None)
        os.environ.update(env)
        tp, tf = timeit(plain), timeit(fused)
        print(f"E {E}  {name:22s} ln_bwd {tp:6.1f} us ({M * E * 14 / tp / 1e3:5.0f} GB/s)   ln_bwd+emit {tf:6.1f} us ({M * E * 16 / tf / 1e3:5.0f} GB/s)")
    print(f"E {E}  cast_scale alone {timeit(cast):6.1f} us")
