"""AutoFormer-structured attention kernels alone at the c3-max shape (for `ncu`): forward, backward rows, backward cols."""
import sys
import torch
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
for _ in range(2):
    out, lse = ops.attention_fwd(qkv, B, h, N, 0.125, tk=tk, tv=tv, idx=(iv, ih, iv, ih), af=(14, 14))
    ops.attention_bwd(qkv, out, lse, dout, B, h, N, 0.125, tk=tk, tv=tv, idx=(iv, ih, iv, ih), af=(14, 14))
torch.cuda.synchronize()
print("ok")
