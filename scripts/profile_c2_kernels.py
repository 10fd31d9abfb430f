"""BASELINE config 2 kernels in isolation (for `ncu`): fused attention + iRPE (grid-product gather) forward /
backward at B 256, 6 heads, N 197, and the rpe_index operator forward / backward at B 256, H 6, L 197, 50 buckets."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cream_b200 import ops
from oracle import rel_index
B, N, h = 256, 197, 6
ids, nb = rel_index.irpe_bucket_ids(rel_index.PRODUCT, 14, 14, 1, 1.9, 3.8, 15.2)
ids = ids.astype(np.int32)
gp = (14,) + ops.irpe_grid_product_structure(ids, 14, 1)
it = ops.irpe_index_table_u8(ids, "cuda")
torch.manual_seed(0)
qkv = ops.empty_bf16(B * N, 3 * 64 * h); qkv.copy_(torch.randn(B * N, 3 * 64 * h, device="cuda"))
dout = ops.empty_bf16(B * N, 64 * h); dout.copy_(torch.randn(B * N, 64 * h, device="cuda"))
tk = ops.new_pack(1, "cuda"); tk.zero_(); tk[0, :nb] = (torch.randn(nb, 64, device="cuda") * 0.3).to(torch.bfloat16)
for g in (gp, None):
    for _ in range(2):
        out, lse = ops.attention_fwd(qkv, B, h, N, 0.125, tk=tk, idx=(it, None, None, None), gp=g)
        ops.attention_bwd(qkv, out, lse, dout, B, h, N, 0.125, tk=tk, idx=(it, None, None, None), gp=g)
x = torch.randn(B, h, N, nb, device="cuda")
idx = torch.from_numpy(ids).cuda()
gy = torch.randn(B, h, N, N, device="cuda")
for _ in range(2):
    y = ops.rpe_index_forward(x, idx)
    gx = torch.zeros_like(x)
    ops.rpe_index_backward(gx, gy, idx)
torch.cuda.synchronize()
print("ok")
