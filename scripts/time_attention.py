"""Stand-alone timing of the fused attention kernels at the c3-max shape (GPU box)."""
import sys, torch
sys.path.insert(0, ".")
from cream_b200 import ops
B, N, h = 128, 197, 7
torch.manual_seed(0)
qkv = ops.empty_bf16(B * N, 3 * 64 * h); qkv.copy_(torch.randn(B * N, 3 * 64 * h, device="cuda"))
dout = ops.empty_bf16(B * N, 64 * h); dout.copy_(torch.randn(B * N, 64 * h, device="cuda"))
iv, ih, _, _ = ops.autoformer_index_tables(N, 14, "cuda")
tk, tv = ops.new_pack(1, "cuda"), ops.new_pack(1, "cuda")
for t in (tk, tv):
    t.zero_(); t[0, :30] = (torch.randn(30, 64, device="cuda") * 0.3).to(torch.bfloat16); t[0, 32:62] = (torch.randn(30, 64, device="cuda") * 0.3).to(torch.bfloat16)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
import os
ONLY = os.environ.get("CREAM_ONLY_STRUCTURED") == "1"
for name, af in ((("structured", (14, 14)),) if ONLY else (("generic", None), ("structured", (14, 14)))):
    out, lse = ops.attention_fwd(qkv, B, h, N, 0.125, tk=tk, tv=tv, idx=(iv, ih, iv, ih), af=af)
    tf = timeit(lambda: ops.attention_fwd(qkv, B, h, N, 0.125, tk=tk, tv=tv, idx=(iv, ih, iv, ih), af=af))
    tb = timeit(lambda: ops.attention_bwd(qkv, out, lse, dout, B, h, N, 0.125, tk=tk, tv=tv, idx=(iv, ih, iv, ih), af=af))
    fl = 4.0 * B * h * N * N * 64 + 2.0 * B * h * N * 64 * 128
    by = 4.0 * B * h * N * 64 * 2
    print(f"{name:10s} fwd {tf:7.1f} us  {fl / tf / 1e6:6.1f} TFLOP/s  {by / tf / 1e3:6.0f} GB/s | bwd (rows+cols) {tb:7.1f} us")
if ONLY:
    sys.exit(0)
# BASELINE config 2 shape: DeiT-S + iRPE product table on keys (B 256, 6 heads): index-table path vs structured path
from oracle import rel_index
import numpy as np
B, h = 256, 6
qkv = ops.empty_bf16(B * N, 3 * 64 * h); qkv.copy_(torch.randn(B * N, 3 * 64 * h, device="cuda"))
dout = ops.empty_bf16(B * N, 64 * h); dout.copy_(torch.randn(B * N, 64 * h, device="cuda"))
ids, nb = rel_index.irpe_bucket_ids(rel_index.PRODUCT, 14, 14, 1, 1.9, 3.8, 15.2)
ids = ids.astype(np.int32)
it = ops.irpe_index_table_u8(ids, "cuda")
gp = (14,) + ops.irpe_grid_product_structure(ids, 14, 1)
tk.zero_(); tk[0, :nb] = (torch.randn(nb, 64, device="cuda") * 0.3).to(torch.bfloat16)
for name, g in (("c2 table", None), ("c2 struct", gp)):
    out, lse = ops.attention_fwd(qkv, B, h, N, 0.125, tk=tk, idx=(it, None, None, None), gp=g)
    tf = timeit(lambda: ops.attention_fwd(qkv, B, h, N, 0.125, tk=tk, idx=(it, None, None, None), gp=g))
    tb = timeit(lambda: ops.attention_bwd(qkv, out, lse, dout, B, h, N, 0.125, tk=tk, idx=(it, None, None, None), gp=g))
    fl = 4.0 * B * h * N * N * 64 + 2.0 * B * h * N * 64 * 64
    by = 4.0 * B * h * N * 64 * 2
    print(f"{name:10s} fwd {tf:7.1f} us  {fl / tf / 1e6:6.1f} TFLOP/s  {by / tf / 1e3:6.0f} GB/s | bwd (rows+cols) {tb:7.1f} us")
