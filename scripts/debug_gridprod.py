"""GPU debug: where does the grid-product backward differ from the index-table backward?"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cream_b200 import ops
from oracle import rel_index
B, N, h = 2, 197, 2
ids, nb = rel_index.irpe_bucket_ids(rel_index.PRODUCT, 14, 14, 1, 1.9, 3.8, 15.2)
ids = ids.astype(np.int32)
gp = (14,) + ops.irpe_grid_product_structure(ids, 14, 1)
it = ops.irpe_index_table_u8(ids, "cuda")
torch.manual_seed(5)
qkv = ops.empty_bf16(B * N, 3 * 64 * h); qkv.copy_(torch.randn(B * N, 3 * 64 * h, device="cuda"))
dout = ops.empty_bf16(B * N, 64 * h); dout.copy_(torch.randn(B * N, 64 * h, device="cuda"))
tk = ops.new_pack(1, "cuda"); tk.zero_(); tk[0, :nb] = (torch.randn(nb, 64, device="cuda") * 0.3).to(torch.bfloat16)
res = {}
for name, g in (("table", None), ("structured", gp)):
    out, lse = ops.attention_fwd(qkv, B, h, N, 0.125, tk=tk, idx=(it, None, None, None), gp=g)
    dqkv, dtk, _, _ = ops.attention_bwd(qkv, out, lse, dout, B, h, N, 0.125, tk=tk, idx=(it, None, None, None), gp=g)
    res[name] = (dqkv.float().view(B, N, 3, h, 64), dtk[0])
a, b = res["table"], res["structured"]
for part, nm in enumerate("qkv"):
    x, y = a[0][:, :, part], b[0][:, :, part]
    bad = ~torch.isfinite(y)
    print(nm, "nan count", int(bad.sum()), "of", y.numel(), "rel err (finite)", float(((x - y)[~bad]).norm() / x[~bad].norm()))
    if bad.any():
        idx = bad.nonzero()
        print("   first bad (b, i, head, d):", idx[:8].tolist(), " rows with nan:", sorted(set(idx[:, 1].tolist()))[:40])
print("dtk nan", int((~torch.isfinite(b[1])).sum()), "rel err", float((a[1] - b[1]).norm() / a[1].norm()))
d = (a[1] - b[1]).abs().sum(1)
print("dtk per-bucket abs diff:", [round(float(v), 4) for v in d[:52]])
print("dtk per-bucket ref norm:", [round(float(v), 3) for v in a[1].abs().sum(1)[:52]])
