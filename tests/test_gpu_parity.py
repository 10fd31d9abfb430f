"""GPU parity tests (B200): every kernel through the C ABI against the oracle on the same
seeded inputs, against the golden fixtures produced by the unmodified reference, and — at
BASELINE.json's full sizes — through size-independent properties.

Tolerances (relative L2 unless stated):
  * integer / index / copy work (rpe_index forward): bit exact;
  * single kernels fed bf16-exact inputs, oracle in fp32 on the SAME bf16 values: 1e-3 .. 4e-3
    (the only differences are fp32 accumulation order and the bf16 rounding of P / outputs);
  * whole sampled subnet in bf16 against the fp32 oracle: logits 1e-2, gradients 3e-2
    (twelve+ chained bf16 GEMMs; BASELINE.json's 1e-3 applies per kernel).
"""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from oracle import rel_index, vit_oracle as vo  # noqa: E402
from tests.helpers import check_summary, rand, rel_err  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cream_b200 import _lib, ops
    _lib.load()
    ops.SHADOWS.clear()
    yield
    torch.cuda.synchronize()


def bf16r(t):
    return t.to(torch.bfloat16).float()


# ------------------------------------------------------------------------------------------
# rpe_index (the reference's native op; its own self-check is rpe_ops/rpe_index.py:59-100)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.float64])
def test_rpe_index_reference_selfcheck(dtype):
    from cream_b200.rpe_ops.rpe_index import RPEIndexFunction
    B, H, L, nb = 128, 32, 50, 50
    torch.manual_seed(0)
    x = torch.randn(B, H, L, nb).to(dtype)
    index = torch.randint(0, nb, (L, L)).int()
    offset = torch.arange(0, L * nb, nb).view(-1, 1)
    x1 = x.cuda().requires_grad_(True)
    y = RPEIndexFunction.apply(x1, index.cuda())
    gt = x.flatten(2)[:, :, (index + offset).flatten()].view(B, H, L, L)
    assert torch.equal(y.detach().cpu(), gt), "forward must be bit exact"
    mask = torch.randn(gt.shape).to(dtype)
    (y * mask.cuda()).sum().backward()
    x2 = x.double().requires_grad_(True)
    (x2.flatten(2)[:, :, (index + offset).flatten()].view(B, H, L, L) * mask.double()).sum().backward()
    tol = {torch.float32: 1e-5, torch.float64: 1e-10, torch.float16: 2e-2, torch.bfloat16: 1e-1}[dtype]
    np.testing.assert_allclose(x1.grad.double().cpu().numpy(), x2.grad.numpy(), atol=tol, rtol=0)


def test_rpe_index_golden_and_strided_input(golden_dir):
    from cream_b200 import ops
    g = np.load(golden_dir / "rpe_index.npz")
    B, H, L, nb = 4, 3, 50, 50
    x = rand((B, H, L, nb), 5)
    idx = torch.from_numpy(np.random.default_rng(6).integers(0, nb, (L, L)).astype(np.int32))
    y = ops.rpe_index_forward(x.cuda(), idx.cuda())
    np.testing.assert_array_equal(y.cpu().numpy(), g["y"])
    # non-contiguous input exactly as irpe.py:639-642 produces it: (H, B, L, nb) transposed view
    xt = x.permute(1, 0, 2, 3).contiguous().cuda().transpose(0, 1)
    assert not xt.is_contiguous()
    np.testing.assert_array_equal(ops.rpe_index_forward(xt, idx.cuda()).cpu().numpy(), g["y"])
    gy = rand((B, H, L, L), 8)
    gx = torch.zeros(B, H, L, nb, device="cuda")
    ops.rpe_index_backward(gx, gy.cuda(), idx.cuda())
    np.testing.assert_allclose(gx.cpu().numpy(), g["gx"], atol=1e-5, rtol=0)


def test_rpe_index_errors_like_reference():
    from cream_b200 import ops
    x = torch.randn(2, 2, 5, 7, device="cuda")
    with pytest.raises(RuntimeError):
        ops.rpe_index_forward(x, torch.zeros(5, 5, dtype=torch.int64, device="cuda"))   # index must be int32
    with pytest.raises(RuntimeError):
        ops.rpe_index_forward(x[0], torch.zeros(5, 5, dtype=torch.int32, device="cuda"))  # 4-D input
    with pytest.raises(RuntimeError):
        ops.rpe_index_forward(x.cpu(), torch.zeros(5, 5, dtype=torch.int32))            # GPU tensors only
    empty = ops.rpe_index_forward(x[:0], torch.zeros(5, 5, dtype=torch.int32, device="cuda"))
    assert empty.shape == (0, 2, 5, 5)


def test_rpe_index_full_size_adjoint():
    """BASELINE config 2 size (B=256, H=6, L=197, nb=50): <fwd(X), G> == <X, bwd(G)>."""
    from cream_b200 import ops
    B, H, L, nb = 256, 6, 197, 50
    ids, n = rel_index.irpe_bucket_ids(rel_index.PRODUCT, 14, 14, 1, 1.9, 3.8, 15.2)
    idx = torch.from_numpy(ids.astype(np.int32)).cuda()
    torch.manual_seed(1)
    x = torch.randn(B, H, L, nb, device="cuda")
    g = torch.randn(B, H, L, L, device="cuda")
    y = ops.rpe_index_forward(x, idx)
    gx = torch.zeros_like(x)
    ops.rpe_index_backward(gx, g, idx)
    lhs = (y.double() * g.double()).sum().item()
    rhs = (x.double() * gx.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0) + 1e-2
    # every output is a copy of an input of its own row
    assert torch.equal(y[3, 2, 17], x[3, 2, 17][idx[17].long()])


# ------------------------------------------------------------------------------------------
# sliced GEMMs
# ------------------------------------------------------------------------------------------
SLICES = [  # (E*, E, heads*, heads, ratio)  supernet-T / S / B corners + ragged dims
    (256, 216, 4, 3, 3.5), (448, 320, 7, 5, 3.0), (448, 448, 7, 7, 4.0), (640, 528, 10, 9, 3.5),
]


@pytest.mark.parametrize("Es,E,hs,h,r", SLICES)
def test_sliced_linear_fwd_dgrad_wgrad(Es, E, hs, h, r):
    from cream_b200 import ops
    ops.SHADOWS.clear()
    M = 394
    ffn_s, ffn = int(Es * 4.0), int(E * r)
    torch.manual_seed(2)
    w = (torch.randn(ffn_s, Es) * 0.05).cuda()
    b = (torch.randn(ffn_s) * 0.1).cuda()
    x = ops.empty_bf16(M, E)
    x.copy_(torch.randn(M, E))
    wr, xr = bf16r(w), x.float()
    y = ops.linear_fwd(x, ops.SHADOWS.get(w), ffn, E, b)
    ref = xr @ wr[:ffn, :E].t() + b[:ffn]
    assert rel_err(y.float().cpu(), ref.cpu()) < 4e-3
    dy = ops.empty_bf16(M, ffn)
    dy.copy_(torch.randn(M, ffn))
    dx = ops.linear_dgrad(dy, ops.SHADOWS.get(w), ffn, E)
    assert rel_err(dx.float().cpu(), (dy.float() @ wr[:ffn, :E]).cpu()) < 4e-3
    dw = torch.zeros_like(w)
    ops.linear_wgrad(dy, x, ffn, E, dw)
    refw = torch.zeros_like(w)
    refw[:ffn, :E] = dy.float().t() @ xr
    assert rel_err(dw.cpu(), refw.cpu()) < 1e-4
    assert float(dw[ffn:].abs().sum()) == 0.0 and float(dw[:, E:].abs().sum()) == 0.0, "zero outside the slice"
    db = torch.zeros_like(b)
    ops.bias_grad(dy, db)
    assert rel_err(db[:ffn].cpu(), dy.float().sum(0).cpu()) < 1e-4 and float(db[ffn:].abs().sum()) == 0.0


@pytest.mark.parametrize("Es,E,hs,h,r", SLICES)
def test_sliced_qkv_interleaved_rows(Es, E, hs, h, r):
    """qkv_super semantics: weight rows {i, i+3, ...} but CONTIGUOUS bias (qkv_super.py:72-83)."""
    from cream_b200 import ops
    ops.SHADOWS.clear()
    M, qd = 394, 64 * h
    torch.manual_seed(3)
    w = (torch.randn(3 * Es, Es) * 0.05).cuda()
    b = (torch.randn(3 * Es) * 0.1).cuda()
    x = ops.empty_bf16(M, E)
    x.copy_(torch.randn(M, E))
    wr = bf16r(w)
    w_s = torch.cat([wr[:, :E][i:3 * qd:3, :] for i in range(3)], dim=0)
    y = ops.qkv_fwd(x, ops.SHADOWS.get(w, qkv=True), h, E, Es, b)
    ref = x.float() @ w_s.t() + b[:3 * qd]
    assert rel_err(y.float().cpu(), ref.cpu()) < 4e-3
    dy = ops.empty_bf16(M, 3 * qd)
    dy.copy_(torch.randn(M, 3 * qd))
    dx = ops.qkv_dgrad(dy, ops.SHADOWS.get(w, qkv=True), h, E, Es)
    assert rel_err(dx.float().cpu(), (dy.float() @ w_s).cpu()) < 4e-3
    dw = torch.zeros_like(w)
    ops.qkv_wgrad(dy, x, h, E, dw)
    g_s = dy.float().t() @ x.float()            # (3qd, E) in [q | k | v] block order
    refw = torch.zeros_like(w)
    for i in range(3):
        refw[i:3 * qd:3, :E] = g_s[i * qd:(i + 1) * qd]
    assert rel_err(dw.cpu(), refw.cpu()) < 1e-4
    assert float(dw[3 * qd:].abs().sum()) == 0.0 and float(dw[:, E:].abs().sum()) == 0.0


def test_gemm_epilogues_gelu_residual():
    from cream_b200 import ops
    from cream_b200._lib import EPI_BF16_DGELU, EPI_BF16_GELU, EPI_F32_RESID
    ops.SHADOWS.clear()
    M, K, Nn, rows_per = 394, 320, 1120, 197
    torch.manual_seed(4)
    w = (torch.randn(Nn, K) * 0.06).cuda()
    b = (torch.randn(Nn) * 0.1).cuda()
    x = ops.empty_bf16(M, K)
    x.copy_(torch.randn(M, K))
    pre = ops.empty_bf16(M, Nn)
    act = ops.linear_fwd(x, ops.SHADOWS.get(w), Nn, K, b, epi=EPI_BF16_GELU, aux=pre)
    ref_pre = x.float() @ bf16r(w).t() + b
    assert rel_err(pre.float().cpu(), ref_pre.cpu()) < 4e-3
    assert rel_err(act.float().cpu(), F.gelu(pre.float()).cpu()) < 4e-3     # exact (erf) GELU in fp32
    dy = ops.empty_bf16(M, K)
    dy.copy_(torch.randn(M, K))
    w2 = (torch.randn(K, Nn) * 0.05).cuda()      # fc2 weight (out=K, in=Nn)
    dh = ops.linear_dgrad(dy, ops.SHADOWS.get(w2), K, Nn, epi=EPI_BF16_DGELU, aux=pre)
    p = pre.float().requires_grad_(True)
    F.gelu(p).backward(dy.float() @ bf16r(w2))
    assert rel_err(dh.float().cpu(), p.grad.cpu()) < 5e-3
    resid = ops.empty_f32(M, K)
    resid.copy_(torch.randn(M, K))
    scale = torch.tensor([0.0, 1.0 / 0.9], device="cuda")
    out = ops.linear_fwd(act, ops.SHADOWS.get(w2), K, Nn, None, epi=EPI_F32_RESID, resid=resid, row_scale=scale,
                         rows_per_scale=rows_per)
    ref = resid + scale.repeat_interleave(rows_per)[:, None] * (act.float() @ bf16r(w2).t())
    assert rel_err(out.cpu(), ref.cpu()) < 1e-3
    assert torch.equal(out[:rows_per], resid[:rows_per]), "a dropped path leaves the residual untouched"


def test_gemm_full_size_linearity():
    """c3-max shapes (25216 tokens): f(a x1 + x2) == a f(x1) + f(x2) within bf16 rounding."""
    from cream_b200 import ops
    from cream_b200._lib import EPI_F32
    ops.SHADOWS.clear()
    M, K, Nn = 128 * 197, 448, 1792
    torch.manual_seed(5)
    w = (torch.randn(Nn, K) * 0.05).cuda()
    x1 = ops.empty_bf16(M, K); x1.copy_(torch.randn(M, K, device="cuda"))
    x2 = ops.empty_bf16(M, K); x2.copy_(torch.randn(M, K, device="cuda"))
    xs = ops.empty_bf16(M, K); xs.copy_(2.0 * x1.float() + x2.float())   # exact in bf16 up to 1 ulp
    sh = ops.SHADOWS.get(w)
    f = lambda x: ops.linear_fwd(x, sh, Nn, K, None, epi=EPI_F32)
    lhs, rhs = f(xs), 2.0 * f(x1) + f(x2)
    assert rel_err(lhs[::97].cpu(), rhs[::97].cpu()) < 5e-3
    # spot check 64 rows against fp32 math on the same bf16 operands
    rows = torch.arange(0, M, M // 64, device="cuda")
    ref = x1[rows].float() @ bf16r(w).t()
    assert rel_err(f(x1)[rows].cpu(), ref.cpu()) < 1e-3


# ------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("E,Es", [(192, 256), (216, 256), (448, 448), (624, 640)])
def test_layernorm_fwd_bwd(E, Es):
    from cream_b200 import ops
    rows = 3 * 197
    torch.manual_seed(6)
    x = ops.empty_f32(rows, E); x.copy_(torch.randn(rows, E) * 2 + 0.5)
    gw = (1 + 0.1 * torch.randn(Es)).cuda()
    gb = (0.1 * torch.randn(Es)).cuda()
    y, mean, rstd = ops.layernorm_fwd(x, gw, gb, 1e-5, E, out_f32=True)
    xr = x.clone().contiguous().requires_grad_(True)
    wr, br = gw.clone().requires_grad_(True), gb.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (E,), wr[:E], br[:E], 1e-5)
    assert rel_err(y.cpu(), ref.detach().cpu()) < 1e-5
    yb, _, _ = ops.layernorm_fwd(x, gw, gb, 1e-5, E)
    assert rel_err(yb.float().cpu(), ref.detach().cpu()) < 4e-3
    dy = ops.empty_f32(rows, E); dy.copy_(torch.randn(rows, E))
    rg = ops.empty_f32(rows, E); rg.copy_(torch.randn(rows, E))
    dgw, dgb = torch.zeros_like(gw), torch.zeros_like(gb)
    dx = ops.layernorm_bwd(dy, x, gw, mean, rstd, E, dgw, dgb, resid_grad=rg)
    ref.backward(dy.clone().contiguous())
    assert rel_err(dx.cpu(), (xr.grad + rg).cpu()) < 1e-4
    assert rel_err(dgw.cpu(), wr.grad.cpu()) < 1e-4 and rel_err(dgb.cpu(), br.grad.cpu()) < 1e-4
    assert float(dgw[E:].abs().sum()) == 0.0


@pytest.mark.parametrize("E,rows,dy_bf16,scaled", [(192, 591, True, True), (448, 3 * 197, True, False), (448, 128 * 197, True, True),
                                                   (624, 1001, False, True), (320, 37, True, True)])
def test_layernorm_bwd_with_fused_cast(E, rows, dy_bf16, scaled):
    """cream_layernorm_bwd_cast == cream_layernorm_bwd followed by cream_cast_scale: dx, the bf16 DropPath-scaled
    copy, the bias-gradient column sums and dgamma / dbeta (pipelined kernel incl. ragged tail tiles; rows = 37 takes
    the two-pass form); dx is also checked against torch autograd."""
    from cream_b200 import ops
    torch.manual_seed(8)
    x = ops.empty_f32(rows, E); x.copy_(torch.randn(rows, E) * 1.5 - 0.3)
    gw, gb = (1 + 0.1 * torch.randn(E)).cuda(), (0.1 * torch.randn(E)).cuda()
    _, mean, rstd = ops.layernorm_fwd(x, gw, gb, 1e-5, E)
    if dy_bf16:
        dy = ops.empty_bf16(rows, E); dy.copy_(torch.randn(rows, E))
    else:
        dy = ops.empty_f32(rows, E); dy.copy_(torch.randn(rows, E))
    rg = ops.empty_f32(rows, E); rg.copy_(torch.randn(rows, E))
    per = 197 if rows % 197 == 0 else rows
    scale = (torch.rand(rows // per, device="cuda") > 0.3).float() / 0.7 if scaled else None
    a_g, a_b, a_bias = torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda")
    b_g, b_b, b_bias = torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda")
    dx_a = ops.layernorm_bwd(dy, x, gw, mean, rstd, E, a_g, a_b, resid_grad=rg)
    bf_a = ops.cast_scale(dx_a, scale, per, dbias=a_bias)
    dx_b, bf_b = ops.layernorm_bwd_cast(dy, x, gw, mean, rstd, E, b_g, b_b, resid_grad=rg, row_scale=scale, rows_per_scale=per,
                                        dbias=b_bias)
    xr = x.clone().contiguous().requires_grad_(True)
    F.layer_norm(xr, (E,), gw, gb, 1e-5).backward(dy.float().contiguous())
    assert rel_err(dx_b.cpu(), (xr.grad + rg).cpu()) < 1e-4
    assert rel_err(dx_b.cpu(), dx_a.cpu()) < 1e-6
    # the bf16 copies may differ by one rounding step where dx differs in its last bits
    assert rel_err(bf_b.float().cpu(), bf_a.float().cpu()) < 1e-3
    assert float((bf_b.float() - bf_a.float()).abs().max()) <= 2 ** -7 * float(bf_a.float().abs().max())
    assert rel_err(b_bias.cpu(), a_bias.cpu()) < 1e-4
    assert rel_err(b_g.cpu(), a_g.cpu()) < 1e-5 and rel_err(b_b.cpu(), a_b.cpu()) < 1e-5


# ------------------------------------------------------------------------------------------
# fused attention
# ------------------------------------------------------------------------------------------
def _qkv(B, N, h, seed, scale=1.0):
    from cream_b200 import ops
    q = ops.empty_bf16(B * N, 3 * 64 * h)
    q.copy_(rand((B * N, 3 * 64 * h), seed, scale))
    return q


def _af_tables(seed, std=0.3):
    return [bf16r(rand((30, 64), seed + i, std)).cuda().requires_grad_(True) for i in range(4)]


@pytest.mark.parametrize("B,N,h,tables", [(2, 197, 3, False), (2, 197, 3, True), (3, 17, 2, True), (1, 50, 1, True),
                                          (2, 65, 2, False)])
def test_attention_autoformer_fwd_bwd(B, N, h, tables):
    from cream_b200.autoformer.functional import AutoformerAttentionFn
    qkv = _qkv(B, N, h, 20).reshape(B, N, -1).requires_grad_(True)
    tabs = _af_tables(30) if tables else []
    out = AutoformerAttentionFn.apply(qkv, h, 0.125, 14, *tabs)
    dout = bf16r(rand((B, N, 64 * h), 40, 1.0)).cuda()
    out.backward(dout.to(out.dtype))
    # oracle on the same bf16 values, fp32 math, CPU
    q_ref = qkv.detach().float().cpu().requires_grad_(True)
    t_ref = [t.detach().cpu().requires_grad_(True) for t in tabs]
    ref = vo.attention_core_autoformer(q_ref.reshape(B, N, 3, h, 64), tuple(t_ref) if tables else None, 14, 0.125)
    ref.backward(dout.cpu())
    assert rel_err(out.float().cpu(), ref.detach()) < 4e-3, "forward"
    assert rel_err(qkv.grad.float().cpu(), q_ref.grad) < 1e-2, "dqkv"
    for name, t, tr in zip(("k_v", "k_h", "v_v", "v_h"), tabs, t_ref):
        assert rel_err(t.grad.cpu(), tr.grad) < 1e-2, f"table grad {name}"


def test_attention_full_size_properties():
    """c3-max size (B=128, 7 heads, N=197): rows of P sum to one -> with v == 1 and zero value
    tables the output is 1; key padding columns contribute nothing."""
    from cream_b200 import ops
    B, N, h = 128, 197, 7
    qkv = ops.empty_bf16(B * N, 3 * 64 * h)
    torch.manual_seed(7)
    qkv.copy_(torch.randn(B * N, 3 * 64 * h, device="cuda"))
    qkv[:, 2 * 64 * h:] = 1.0
    iv, ih, _, _ = ops.autoformer_index_tables(N, 14, "cuda")
    tk = ops.new_pack(1, "cuda")
    tk.copy_(torch.randn(1, 64, 64, device="cuda") * 0.2)
    tv = torch.zeros_like(tk)
    out, lse = ops.attention_fwd(qkv, B, h, N, 0.125, tk=tk, tv=tv, idx=(iv, ih, iv, ih))
    assert torch.isfinite(lse).all()
    assert (out.float() - 1.0).abs().max().item() < 1e-2


IRPE_GPU_CASES = {
    # name: (rpe_on, mode, shared_head, method, heads, grid)
    "k_ctx_shared": ("k", "contextual", True, "product", 2, 14),
    "kv_ctx_perhead": ("kv", "contextual", False, "product", 2, 7),
    "k_bias": ("k", "bias", False, "euc", 2, 7),
    "k_ctx_quant": ("k", "contextual", True, "quant", 1, 5),
    "kv_ctx_cross": ("kv", "contextual", True, "cross", 2, 7),
    "k_ctx_cross_perhead": ("k", "contextual", False, "cross", 2, 14),
    "qkv_ctx_perhead": ("qkv", "contextual", False, "product", 2, 7),     # deit_*_shared_qkv kind
    "qk_ctx_shared": ("qk", "contextual", True, "product", 2, 14),
    "qk_bias": ("qk", "bias", False, "euc", 2, 7),
    "q_ctx_cross": ("q", "contextual", True, "cross", 1, 5),
}


@pytest.mark.parametrize("name", list(IRPE_GPU_CASES))
def test_irpe_attention_module(name):
    from cream_b200.irpe_attention import RPEAttention
    rpe_on, mode, shared, method, heads, grid = IRPE_GPU_CASES[name]
    C, B, N = 64 * heads, 2, grid * grid + 1
    m = RPEAttention(C, num_heads=heads, qkv_bias=True, rpe_on=rpe_on, method=method, mode=mode,
                     shared_head=shared).cuda()
    seed = 200
    with torch.no_grad():
        for pn, p in m.named_parameters():
            seed += 1
            p.copy_(bf16r(rand(tuple(p.shape), seed, 0.3 if "lookup" in pn else 0.08)))
    x = bf16r(rand((B, N, C), 199)).cuda().requires_grad_(True)
    gy = bf16r(rand((B, N, C), 198)).cuda()
    y = m(x)
    y.backward(gy.to(y.dtype))
    # oracle (fp32, CPU) with the same parameters
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.named_parameters()}
    xr = x.detach().cpu().clone().requires_grad_(True)
    if method == "cross":
        ids = tuple(rel_index.irpe_bucket_ids(mm, grid, grid, 1, 1.9, 3.8, 15.2)[0]
                    for mm in (rel_index.CROSS_ROWS, rel_index.CROSS_COLS))
        for a, b in zip(ids, m.bucket_ids(N)):
            np.testing.assert_array_equal(a, b)                        # library ids == oracle ids (bit exact)
    else:
        mid = {"product": rel_index.PRODUCT, "euc": rel_index.EUCLIDEAN, "quant": rel_index.QUANT}[method]
        ids, nb = rel_index.irpe_bucket_ids(mid, grid, grid, 1, 1.9, 3.8, 15.2)
        np.testing.assert_array_equal(ids, m.bucket_ids(N))            # library ids == oracle ids (bit exact)

    def tab(w):
        hits = [v for k, v in P.items() if k.startswith(f"rpe_{w}.")]     # cross: rp_rows, rp_cols
        return None if not hits else (hits[0] if len(hits) == 1 else tuple(hits))
    ref = vo.rpe_attention(xr, P["qkv.weight"], P["qkv.bias"], P["proj.weight"], P["proj.bias"], heads, ids,
                           rpe_q=tab("q"), rpe_k=tab("k"), rpe_v=tab("v"), mode=mode)
    ref.backward(gy.cpu())
    assert rel_err(y.float().cpu(), ref.detach()) < 1e-2
    assert rel_err(x.grad.float().cpu(), xr.grad) < 2e-2
    for pn, p in m.named_parameters():
        assert rel_err(p.grad.float().cpu(), P[pn].grad) < 2e-2, pn


@pytest.mark.parametrize("name", ["k_ctx_shared", "kv_ctx_cross", "qkv_ctx_perhead"])   # qk_bias fixture: head_dim 32
def test_irpe_attention_golden_k_ctx_shared(golden_dir, name):
    """BASELINE config-2 kind (contextual product on keys, shared head), and the cross method on
    keys and values, against the fixtures written by the reference's own RPEAttention."""
    from cream_b200.irpe_attention import RPEAttention
    from make_golden import IRPE_CASES
    g = np.load(golden_dir / "irpe_attention.npz")
    rpe_on, mode, shared, method, C, heads, grid = IRPE_CASES[name]
    m = RPEAttention(C, num_heads=heads, qkv_bias=True, rpe_on=rpe_on, method=method,
                     mode="bias" if mode == "bias" else "contextual", shared_head=shared).cuda()
    shapes = {k[len(name) + 7:]: tuple(int(v) for v in g[k]) for k in g.files if k.startswith(name + "_shape_")}
    seed = 100
    with torch.no_grad():
        for pn, shape in shapes.items():
            seed += 1
            dict(m.named_parameters())[pn].copy_(rand(shape, seed, 0.3 if "lookup" in pn else 0.08))
    N = grid * grid + 1
    x = rand((2, N, C), 99).cuda().requires_grad_(True)
    y = m(x)
    y.backward(rand((2, N, C), 98).cuda().to(y.dtype))
    check_summary(g, f"{name}_y", y.float(), 1.5e-2)
    check_summary(g, f"{name}_gx", x.grad.float(), 3e-2)
    for pn, p in m.named_parameters():
        check_summary(g, f"{name}_grad_{pn}", p.grad.float(), 3e-2, what=pn)


# ------------------------------------------------------------------------------------------
# whole sampled subnet: fused engine and module-by-module path vs oracle + golden
# ------------------------------------------------------------------------------------------
def _build(spec, fused):
    from cream_b200.autoformer.model.supernet_transformer import Vision_TransformerSuper
    net = Vision_TransformerSuper(img_size=spec.img_size, patch_size=spec.patch_size, embed_dim=spec.embed_dim,
                                  depth=spec.depth, num_heads=spec.num_heads, mlp_ratio=spec.mlp_ratio,
                                  qkv_bias=True, drop_rate=0.0, drop_path_rate=0.0, gp=True,
                                  num_classes=spec.num_classes, max_relative_position=14, relative_position=True,
                                  change_qkv=True, abs_pos=True, fused=fused)
    net.load_state_dict(vo.init_params(spec, seed=7))
    return net.cuda().train()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", ["micro17", "micro197"])
def test_supernet_vs_golden_and_oracle(golden_dir, name, fused):
    from cream_b200 import ops
    from make_golden import MICRO_SPECS
    ops.SHADOWS.clear()
    g = np.load(golden_dir / "supernet_micro.npz")
    spec, batch, configs = MICRO_SPECS[name]
    net = _build(spec, fused)
    images = rand((batch, 3, spec.img_size, spec.img_size), seed=11).cuda()
    targets = torch.from_numpy(np.random.default_rng(13).integers(0, spec.num_classes, batch)).cuda()
    for ci, cfg in enumerate(configs):
        net.zero_grad(set_to_none=True)
        net.set_sample_config(cfg)
        assert net.get_sampled_params_numel(cfg) == int(g[f"{name}_c{ci}_numel"])
        logits = net(images)
        loss = F.cross_entropy(logits.float(), targets)
        loss.backward()
        key = f"{name}_c{ci}"
        e = rel_err(logits.float().cpu(), g[key + "_logits"])
        assert e < 1e-2, f"{key} logits rel err {e:.3e}"
        assert abs(loss.item() - float(g[key + "_loss"])) < 2e-2
        none = set(g[key + "_none"].tolist())
        worst = 0.0
        for pn, p in net.named_parameters():
            if pn in none:
                assert p.grad is None, f"{pn}: identity-layer parameters must get no gradient"
            else:
                assert p.grad is not None, pn
                worst = max(worst, check_summary(g, f"{key}_grad_{pn}", p.grad.float(), 3e-2, what=f"{key} grad {pn}"))
        print(f"{key} fused={fused}: logits rel err {e:.2e}, worst grad rel err {worst:.2e}")


def test_supernet_fused_matches_module_path():
    from cream_b200 import ops
    from make_golden import MICRO_SPECS
    ops.SHADOWS.clear()
    spec, batch, configs = MICRO_SPECS["micro197"]
    a, b = _build(spec, True), _build(spec, False)
    images = rand((batch, 3, spec.img_size, spec.img_size), seed=21).cuda()
    a.set_sample_config(configs[1]); b.set_sample_config(configs[1])
    ya, yb = a(images), b(images)
    assert rel_err(ya.float().cpu(), yb.float().cpu()) < 1e-2
    a.eval()
    with torch.no_grad():
        ye = a(images)
    assert rel_err(ye.float().cpu(), ya.float().detach().cpu()) < 1e-3, "eval forward == train forward (no dropout)"


@pytest.mark.parametrize("size", ["S", "T", "B"])
def test_supernet_s_random_configs_finite(size):
    """BASELINE config 3 / 1 / 5 shapes: supernet-S (and T, B), the engine's own sample_configs
    sequence (random.seed(epoch), supernet_engine.py:13-24,36), bs 16: finite loss and grads,
    unsampled slices exactly zero, identity layers without grad."""
    import random
    from cream_b200 import ops
    ops.SHADOWS.clear()
    spec = {"S": vo.SUPERNET_S, "T": vo.SUPERNET_T, "B": vo.SUPERNET_B}[size]
    net = _build(spec, True)
    rnd = random.Random(0)
    images = torch.randn(16, 3, 224, 224, device="cuda")
    targets = torch.randint(0, 1000, (16,), device="cuda")
    for step in range(3):
        cfg = vo.sample_configs(vo.SEARCH_SPACE[size], rnd)
        net.zero_grad(set_to_none=True)
        net.set_sample_config(cfg)
        loss = F.cross_entropy(net(images).float(), targets)
        loss.backward()
        assert torch.isfinite(loss)
        E, L = cfg["embed_dim"][0], cfg["layer_num"]
        for pn, p in net.named_parameters():
            if pn.startswith("blocks.") and int(pn.split(".")[1]) >= L:
                assert p.grad is None
                continue
            assert torch.isfinite(p.grad).all(), pn
        gq = net.blocks[0].attn.qkv.weight.grad
        qd = 64 * cfg["num_heads"][0]
        assert float(gq[3 * qd:].abs().sum()) == 0.0 and float(gq[:, E:].abs().sum()) == 0.0
        assert float(net.pos_embed.grad[..., E:].abs().sum()) == 0.0


def test_cpu_tensors_fail_loudly():
    from make_golden import MICRO_SPECS
    from cream_b200.autoformer.model.supernet_transformer import Vision_TransformerSuper
    spec, batch, configs = MICRO_SPECS["micro17"]
    net = Vision_TransformerSuper(img_size=64, embed_dim=128, depth=3, num_heads=2, mlp_ratio=4.0, qkv_bias=True,
                                  gp=True, relative_position=True, change_qkv=True)
    net.set_sample_config(configs[0])
    with pytest.raises(RuntimeError):
        net(torch.randn(1, 3, 64, 64))


def test_attention_structured_matches_generic_full_size():
    """c3-max size: the AutoFormer-structured gather (registers) and the generic gather (index
    tables through shared memory) must agree, forward and backward."""
    from cream_b200 import ops
    B, N, h = 16, 197, 7
    torch.manual_seed(8)
    qkv = ops.empty_bf16(B * N, 3 * 64 * h)
    qkv.copy_(torch.randn(B * N, 3 * 64 * h, device="cuda"))
    dout = ops.empty_bf16(B * N, 64 * h)
    dout.copy_(torch.randn(B * N, 64 * h, device="cuda"))
    iv, ih, _, _ = ops.autoformer_index_tables(N, 14, "cuda")
    tk, tv = ops.new_pack(1, "cuda"), ops.new_pack(1, "cuda")
    for t in (tk, tv):
        t.zero_()
        t[0, :30] = (torch.randn(30, 64, device="cuda") * 0.3).to(torch.bfloat16)
        t[0, 32:62] = (torch.randn(30, 64, device="cuda") * 0.3).to(torch.bfloat16)
    res = {}
    for name, af in (("generic", None), ("structured", (14, 14))):
        out, lse = ops.attention_fwd(qkv, B, h, N, 0.125, tk=tk, tv=tv, idx=(iv, ih, iv, ih), af=af)
        dqkv, dtk, dtv, _ = ops.attention_bwd(qkv, out, lse, dout, B, h, N, 0.125, tk=tk, tv=tv,
                                              idx=(iv, ih, iv, ih), af=af)
        res[name] = (out.float(), lse, dqkv.float(), dtk, dtv)
    for a, b, tol in zip(res["generic"], res["structured"], (2e-3, 1e-4, 1e-2, 1e-2, 1e-2)):
        assert rel_err(a.cpu(), b.cpu()) < tol


@pytest.mark.parametrize("B,N,h,kind", [(2, 77, 2, "causal"), (3, 50, 2, "per_head"), (2, 197, 3, "full")])
def test_attention_dense_logit_term(B, N, h, kind):
    """The dense additive logit term of cream_attn_desc: a causal -inf mask broadcast over batch and
    heads (TinyCLIP text tower, open_clip/model.py:756-762), a per-head (H, N, N) bias (TinyViT
    attention_biases[:, idxs], tiny_vit.py:281-283) and a full (B, H, N, N) term, against a plain
    fp32 PyTorch attention on the same bf16-rounded inputs; dS against autograd."""
    from cream_b200 import ops
    torch.manual_seed(21)
    qkv = bf16r(torch.randn(B, N, 3, h, 64) * 0.7)
    dout = bf16r(torch.randn(B, N, h, 64))
    if kind == "causal":
        dense = torch.full((1, 1, N, N), float("-inf")).triu_(1)
    elif kind == "per_head":
        dense = torch.randn(1, h, N, N)
    else:
        dense = torch.randn(B, h, N, N)
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3).clone().requires_grad_(True) for i in range(3)]
    dref = dense.clone().requires_grad_(kind != "causal")
    att = (torch.matmul(q, k.transpose(-1, -2)) * 0.125 + dref).softmax(-1)
    ref = torch.matmul(att, v).permute(0, 2, 1, 3)                       # (B, N, h, 64)
    ref.backward(dout)
    g = ops.empty_bf16(B * N, 3 * 64 * h)
    g.copy_(qkv.reshape(B * N, -1).cuda())
    do = ops.empty_bf16(B * N, 64 * h)
    do.copy_(dout.reshape(B * N, -1).cuda())
    dd = dense.cuda().contiguous()
    out, lse = ops.attention_fwd(g, B, h, N, 0.125, dense=dd)
    ddense = torch.empty((B, h, N, N), dtype=torch.float32, device="cuda")
    dqkv, _, _, _ = ops.attention_bwd(g, out, lse, do, B, h, N, 0.125, dense=dd, ddense=ddense)
    assert rel_err(out.float().cpu().view(B, N, h, 64), ref.detach()) < 4e-3
    got = dqkv.float().cpu().view(B, N, 3, h, 64)
    for i, t in enumerate((q, k, v)):
        assert rel_err(got[:, :, i].permute(0, 2, 1, 3), t.grad) < 1e-2
    if kind == "causal":
        assert float(ddense.triu(1).abs().max()) == 0.0              # masked logits receive no gradient
    else:
        want = dref.grad if kind == "full" else dref.grad            # autograd already reduced broadcast dims
        have = ddense.cpu() if kind == "full" else ddense.cpu().sum(0, keepdim=True)
        assert rel_err(have, want) < 1e-2


@pytest.mark.parametrize("L,B,heads,causal,prune", [(50, 4, 2, False, False), (77, 3, 2, True, False), (77, 2, 2, True, True)])
def test_clip_attention_module(L, B, heads, causal, prune):
    """TinyCLIP ResidualAttentionBlock.attention mirror (LND layout, in_proj chunking, causal additive
    mask, head_z / hidden_z multipliers) against the oracle restatement of model.py:238-283."""
    from cream_b200.clip_attention import ClipAttention
    E = 64 * heads
    torch.manual_seed(31)
    m = ClipAttention(E, heads).cuda()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(bf16r(torch.randn(p.shape) * 0.08))
    x = bf16r(torch.randn(L, B, E)).cuda().requires_grad_(True)
    gy = bf16r(torch.randn(L, B, E)).cuda()
    mask = torch.full((L, L), float("-inf")).triu_(1) if causal else None
    head_z = bf16r(torch.rand(1, heads, 1, 1)) if prune else None
    hidden_z = bf16r(torch.rand(E)) if prune else None
    y = m(x, mask.cuda() if causal else None, head_z=head_z.cuda() if prune else None,
          hidden_z=hidden_z.cuda() if prune else None)
    y.backward(gy.to(y.dtype))
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.named_parameters()}
    xr = x.detach().cpu().clone().requires_grad_(True)
    ref = vo.clip_attention(xr, P["in_proj_weight"], P["in_proj_bias"], P["out_proj.weight"], P["out_proj.bias"],
                            heads, mask, head_z, hidden_z)
    ref.backward(gy.cpu())
    assert rel_err(y.float().cpu(), ref.detach()) < 1e-2
    assert rel_err(x.grad.float().cpu(), xr.grad) < 2e-2
    for pn, p in m.named_parameters():
        assert rel_err(p.grad.float().cpu(), P[pn].grad) < 2e-2, pn


@pytest.mark.parametrize("name", ["win7", "full14"])
def test_tinyvit_attention_module(golden_dir, name):
    """TinyViT Attention mirror (LayerNorm, qkv, per-head bias gather, softmax, proj; head_dim 32 run
    zero-padded to 64) against the fixture written by the reference's own module."""
    from cream_b200.tinyvit_attention import Attention
    from make_golden import TINYVIT_CASES
    g = np.load(golden_dir / "tinyvit_attention.npz")
    dim, key_dim, heads, ratio, res, B = TINYVIT_CASES[name]
    m = Attention(dim, key_dim, heads, attn_ratio=ratio, resolution=res).cuda().train()
    np.testing.assert_array_equal(m.attention_bias_idxs.cpu().numpy(), g[f"{name}_idxs"])
    shapes = {k[len(name) + 7:]: tuple(int(v) for v in g[k]) for k in g.files if k.startswith(name + "_shape_")}
    seed = 300
    with torch.no_grad():
        for pn, shape in shapes.items():
            seed += 1
            p = dict(m.named_parameters())[pn]
            assert tuple(p.shape) == shape, pn
            if pn == "norm.weight":
                p.copy_(1.0 + rand(shape, seed, 0.1))
            else:
                p.copy_(rand(shape, seed, 0.5 if "attention_biases" in pn else 0.08))
    N = res[0] * res[1]
    x = rand((B, N, dim), 299).cuda().requires_grad_(True)
    y = m(x)
    y.backward(rand((B, N, dim), 298).cuda().to(y.dtype))
    check_summary(g, f"{name}_y", y.float(), 1.5e-2)
    check_summary(g, f"{name}_gx", x.grad.float(), 3e-2)
    for pn, p in m.named_parameters():
        check_summary(g, f"{name}_grad_{pn}", p.grad.float(), 3e-2, what=pn)


@pytest.mark.parametrize("B,N,h", [(3, 77, 2), (2, 16, 1), (2, 197, 3)])
def test_attention_causal_flag_equals_the_dense_causal_mask(B, N, h):
    """cream_attn_desc.causal computes the text tower's mask (open_clip/model.py:756-762) from the coordinates; it
    must give exactly what the same mask passed as the dense additive term gives, forward and backward."""
    from cream_b200 import ops
    qkv = _qkv(B, N, h, 71)
    dout = ops.empty_bf16(B * N, 64 * h)
    dout.copy_(rand((B * N, 64 * h), 72))
    mask = torch.full((N, N), float("-inf")).triu_(1).reshape(1, 1, N, N).cuda()
    o1, l1 = ops.attention_fwd(qkv, B, h, N, 0.125, dense=mask)
    o2, l2 = ops.attention_fwd(qkv, B, h, N, 0.125, causal=True)
    assert torch.equal(o1, o2) and torch.equal(l1, l2)
    d1 = ops.attention_bwd(qkv, o1, l1, dout, B, h, N, 0.125, dense=mask)[0]
    d2 = ops.attention_bwd(qkv, o2, l2, dout, B, h, N, 0.125, causal=True)[0]
    assert torch.isfinite(d2.float()).all()
    assert rel_err(d2.float().cpu(), d1.float().cpu()) < 1e-5


def test_attention_packed_items_equal_separate_items():
    """cream_attn_desc.block_len: two 50-token items per sequence (block-diagonal visibility) against the same items
    as separate batch entries - forward, lse and dqkv; with and without the causal mask (TinyCLIP towers)."""
    from cream_b200 import ops
    B, N, h = 6, 50, 3
    torch.manual_seed(21)
    qkv = ops.empty_bf16(B * N, 3 * 64 * h); qkv.copy_(torch.randn(B * N, 3 * 64 * h, device="cuda"))
    dout = ops.empty_bf16(B * N, 64 * h); dout.copy_(torch.randn(B * N, 64 * h, device="cuda"))
    for causal in (False, True):
        o1, l1 = ops.attention_fwd(qkv, B, h, N, 0.125, causal=causal)
        o2, l2 = ops.attention_fwd(qkv, B // 2, h, 2 * N, 0.125, causal=causal, block=N)
        assert rel_err(o2, o1) < 1e-3, causal        # bf16 outputs; the row sums run over other chunk boundaries
        # lse is (B, H, N): the packed call returns (B/2, H, 2N) = two items side by side per head
        l2u = l2.view(B // 2, h, 2, N).permute(0, 2, 1, 3).reshape(B, h, N)
        assert rel_err(l2u, l1) < 1e-5, causal
        g1 = ops.attention_bwd(qkv, o1, l1, dout, B, h, N, 0.125, causal=causal)[0]
        g2 = ops.attention_bwd(qkv, o2, l2, dout, B // 2, h, 2 * N, 0.125, causal=causal, block=N)[0]
        assert rel_err(g2, g1) < 2e-3, causal
