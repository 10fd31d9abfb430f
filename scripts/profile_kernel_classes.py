"""Every kernel class of the path once, at BASELINE shapes, between cudaProfilerStart/Stop (for
`ncu --profile-from-start off`): the six GEMM epilogues in both tilings, the three attention gather paths forward and
backward, LayerNorm forward / backward, cast, bias-gradient column sums, rpe_index forward / backward.
Prints the launch order so the summary lines can be labelled."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cream_b200 import _lib, ops
from cream_b200._lib import EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_F32, EPI_F32_ATOMIC, EPI_F32_RESID
from oracle import rel_index

dev = "cuda"
torch.manual_seed(0)
ORDER = []


def once(name, fn):
    fn()                                  # warm: tensor maps, function attributes
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ORDER.append(name)


def bf(rows, cols):
    t = ops.empty_bf16(rows, cols, dev)
    t.copy_(torch.randn(rows, cols, device=dev) * 0.5)
    return t


# ---- GEMMs at the supernet-S max subnet: M = 128 * 197, E = 448, ffn = 1792, 7 heads ----
M, E, F, H = 128 * 197, 448, 1792, 7
x, hid = bf(M, E), bf(M, F)
w1, w2, wq = bf(F, E), bf(E, F), bf(3 * E, E)
b1, b2, bq = (torch.randn(n, device=dev) for n in (F, E, 3 * E))
resid = torch.randn(M, E, device=dev)
for pair in (1, 2):
    tag = "single" if pair == 1 else "pair"
    out_q = ops.empty_bf16(M, 3 * E, dev)
    once(f"gemm epi0 bf16+bias  qkv fwd {M}x({3}x{E})x{E} {tag}",
         lambda: ops.gemm(M, E, E, x, x.stride(0), wq, wq.stride(0), out_q, out_q.stride(0), EPI_BF16, groups=3,
                          b_group_rows=E, out_g_col=E, bias=bq, cta_pair=pair))
    act, hpre = ops.empty_bf16(M, F, dev), ops.empty_bf16(M, F, dev)
    once(f"gemm epi1 gelu       fc1 fwd {M}x{F}x{E} {tag}",
         lambda: ops.gemm(M, F, E, x, x.stride(0), w1, w1.stride(0), act, act.stride(0), EPI_BF16_GELU, aux=hpre,
                          ldaux=hpre.stride(0), bias=b1, cta_pair=pair))
    y = ops.empty_f32(M, E, dev)
    once(f"gemm epi2 resid      fc2 fwd {M}x{E}x{F} {tag}",
         lambda: ops.gemm(M, E, F, hid, hid.stride(0), w2, w2.stride(0), y, y.stride(0), EPI_F32_RESID, bias=b2, resid=resid,
                          ldr=resid.stride(0), cta_pair=pair))
    dh = ops.empty_bf16(M, F, dev)
    once(f"gemm epi3 dgelu      fc2 dgrad {M}x{F}x{E} {tag}",
         lambda: ops.gemm(M, F, E, x, x.stride(0), w2, w2.stride(0), dh, dh.stride(0), EPI_BF16_DGELU, b_mn=1, aux=hpre,
                          ldaux=hpre.stride(0), cta_pair=pair))
    dw = torch.zeros(F, E, device=dev)
    once(f"gemm epi4 wgrad      fc1 wgrad {F}x{E}x{M} {tag}",
         lambda: ops.gemm(F, E, M, hid, hid.stride(0), x, x.stride(0), dw, dw.stride(0), EPI_F32_ATOMIC, a_mn=1, b_mn=1,
                          cta_pair=pair))
pooled, wh, bh = bf(128, E), bf(1000, E), torch.randn(1000, device=dev)
logits = ops.empty_f32(128, 1000, dev)
once("gemm epi5 f32+bias   head 128x1000x448 single",
     lambda: ops.gemm(128, 1000, E, pooled, pooled.stride(0), wh, wh.stride(0), logits, logits.stride(0), EPI_F32, bias=bh))

# ---- attention: AutoFormer structured (c3 max), iRPE grid-product and index-table (c2) ----
B, N = 128, 197
qkv, dout = bf(B * N, 3 * 64 * H), bf(B * N, 64 * H)
iv, ih, _, _ = ops.autoformer_index_tables(N, 14, dev)
tk, tv = ops.new_pack(1, dev), ops.new_pack(1, dev)
for t in (tk, tv):
    t.zero_()
    t[0, :30] = (torch.randn(30, 64, device=dev) * 0.3).to(torch.bfloat16)
    t[0, 32:62] = (torch.randn(30, 64, device=dev) * 0.3).to(torch.bfloat16)
st = {}
once("attention fwd  AutoFormer structured B128 h7 N197",
     lambda: st.update(zip(("o", "l"), ops.attention_fwd(qkv, B, H, N, 0.125, tk=tk, tv=tv, idx=(iv, ih, iv, ih), af=(14, 14)))))
once("attention bwd  AutoFormer structured (rows, cols)",
     lambda: ops.attention_bwd(qkv, st["o"], st["l"], dout, B, H, N, 0.125, tk=tk, tv=tv, idx=(iv, ih, iv, ih), af=(14, 14)))
B2, H2 = 256, 6
ids, nb = rel_index.irpe_bucket_ids(rel_index.PRODUCT, 14, 14, 1, 1.9, 3.8, 15.2)
ids = ids.astype(np.int32)
gp = (14,) + ops.irpe_grid_product_structure(ids, 14, 1)
it = ops.irpe_index_table_u8(ids, dev)
qkv2, dout2 = bf(B2 * N, 3 * 64 * H2), bf(B2 * N, 64 * H2)
tk2 = ops.new_pack(1, dev)
tk2.zero_()
tk2[0, :nb] = (torch.randn(nb, 64, device=dev) * 0.3).to(torch.bfloat16)
for name, g in (("iRPE grid-product", gp), ("index-table gather", None)):
    once(f"attention fwd  {name} B256 h6 N197 (config 2)",
         lambda: st.update(zip(("o2", "l2"), ops.attention_fwd(qkv2, B2, H2, N, 0.125, tk=tk2, idx=(it, None, None, None), gp=g))))
    once(f"attention bwd  {name} (rows, cols)",
         lambda: ops.attention_bwd(qkv2, st["o2"], st["l2"], dout2, B2, H2, N, 0.125, tk=tk2, idx=(it, None, None, None), gp=g))

# ---- byte movers at M x E ----
xf = torch.randn(M, E, device=dev)
gam, bet = torch.randn(E, device=dev), torch.randn(E, device=dev)
once("layernorm fwd (bf16 out)", lambda: st.update(zip(("ln", "mu", "rs"), ops.layernorm_fwd(xf, gam, bet, 1e-5, E))))
dg, db = torch.zeros(E, device=dev), torch.zeros(E, device=dev)
once("layernorm bwd (+ residual gradient)", lambda: ops.layernorm_bwd(x, xf, gam, st["mu"], st["rs"], E, dg, db, resid_grad=resid))
dbias = torch.zeros(E, device=dev)
once("cast_scale (+ bias gradient)", lambda: ops.cast_scale(resid, dbias=dbias))
dbf = torch.zeros(F, device=dev)
once("bias gradient column sums", lambda: ops.bias_grad(hid, dbf))

# ---- rpe_index at B 256, H 6, L 197, 50 buckets ----
xr = torch.randn(B2, H2, N, nb, device=dev)
idx = torch.from_numpy(ids).to(dev)
gy = torch.randn(B2, H2, N, N, device=dev)
once("rpe_index fwd B256 H6 L197", lambda: ops.rpe_index_forward(xr, idx))
gx = torch.zeros_like(xr)
once("rpe_index bwd B256 H6 L197", lambda: ops.rpe_index_backward(gx, gy, idx))
print("ORDER:")
for i, n in enumerate(ORDER):
    print(f"  {i:2d}  {n}")
