"""Tensor-level wrappers over the C ABI (torch is used for device memory and streams only).

All functions launch on the current CUDA stream and never synchronise.  Activations are
2-D bf16 tensors `(rows, cols)` with `stride(1) == 1` and a row pitch that is a multiple
of 8 elements (see `empty_bf16`).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional

import numpy as np
import weakref

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_F32, EPI_F32_ATOMIC, EPI_F32_RESID,
                   AttnDesc, GemmDesc, check)

HEAD_DIM = 64
NB_PACK = 64
PROFILE = None   # set to a list to record (kind, start_event, end_event, flops, bytes) per launch


_tls = threading.local()


def _profiled(kind: str, flops: float, nbytes: float, launch):
    """Run `launch()`; when PROFILE is a list, bracket it with CUDA events on the current stream."""
    if PROFILE is None:
        return launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = launch()
    e1.record()
    PROFILE.append((kind, e0, e1, flops, nbytes))
    return r


def _stream() -> int:
    """Current torch stream handle; also binds this host thread (autograd workers included) to
    torch's current device inside the library's own CUDA runtime instance."""
    dev = torch.cuda.current_device()
    if getattr(_tls, "dev", None) != dev:
        check(_lib.load().cream_bind_device(dev), "cream_bind_device", kernels=0)
        _tls.dev = dev
    return torch.cuda.current_stream().cuda_stream


def _p(t) -> Optional[int]:
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def empty_bf16(rows: int, cols: int, device=None, zero: bool = False) -> torch.Tensor:
    """(rows, cols) bf16 view over a buffer whose row pitch is a multiple of 8 elements."""
    ld = round_up(max(cols, 1), 8)
    buf = (torch.zeros if zero else torch.empty)((rows, ld), dtype=torch.bfloat16, device=device or "cuda")
    return buf[:, :cols]


def empty_f32(rows: int, cols: int, device=None, zero: bool = False) -> torch.Tensor:
    ld = round_up(max(cols, 1), 4)
    buf = (torch.zeros if zero else torch.empty)((rows, ld), dtype=torch.float32, device=device or "cuda")
    return buf[:, :cols]


def as_bf16_2d(x: torch.Tensor) -> torch.Tensor:
    """Any (..., C) float tensor -> (rows, C) bf16 with an 8-aligned pitch (copy only if needed)."""
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype == torch.bfloat16 and x2.stride(1) == 1 and x2.stride(0) % 8 == 0 and x2.data_ptr() % 16 == 0:
        return x2
    out = empty_bf16(x2.shape[0], x2.shape[1], x2.device)
    out.copy_(x2)
    return out


def _check_2d(t: torch.Tensor, dtype, name: str, mult: int):
    assert t.is_cuda and t.dtype == dtype and t.dim() == 2 and t.stride(1) == 1, f"{name}: bad tensor"
    assert t.stride(0) % mult == 0, f"{name}: row pitch {t.stride(0)} not a multiple of {mult}"


# --------------------------------------------------------------------------------------------
# weight shadows
# --------------------------------------------------------------------------------------------
class ShadowCache:
    """bf16 shadows of fp32 master weights, refreshed when the parameter's version changes
    (i.e. once per optimizer step).  QKV weights are stored de-interleaved.

    The staleness tag is (tensor version, storage, shape).  In-place updates that go through `.data`
    (`p.data.add_()`, some third-party optimizers, manual EMA / weight surgery) do NOT bump the version
    counter: after such an update call `SHADOWS.invalidate()` (or `invalidate(p)`), e.g. from an
    optimizer post-step hook; `load_state_dict` and torch's own optimizers bump the version and need
    nothing.

    An entry keeps its parameter's STORAGE alive.  Without that, a model that is freed and another one
    created afterwards can put a same-shaped parameter with the same version counter at the same address
    (the caching allocator recycles blocks, and so does the allocator of the storage objects), and the new
    model would silently run on the old model's bf16 weights (seen once a process had built enough models:
    tests/test_gpu_clip.py after the parity and native suites).  Entries whose parameter object has been
    collected are dropped - releasing storage and shadow - every 1024 lookups and by `clear()`."""

    def __init__(self):
        self._store = {}      # (data_ptr, qkv) -> (shadow, tag | None, storage, weakref to the parameter | None)
        self._gets = 0

    def invalidate(self, w: Optional[torch.Tensor] = None) -> None:
        """Force a re-cast on next use: of every shadow, or only of parameter `w`."""
        if w is None:
            self._store = {k: (e[0], None) + tuple(e[2:]) for k, e in self._store.items()}
        else:
            for k in [(w.data_ptr(), False), (w.data_ptr(), True)]:
                if k in self._store:
                    e = self._store[k]
                    self._store[k] = (e[0], None) + tuple(e[2:])

    def _purge(self) -> None:
        dead = [k for k, e in self._store.items() if e[3] is not None and e[3]() is None]
        for k in dead:
            del self._store[k]

    def get(self, w: torch.Tensor, qkv: bool = False) -> torch.Tensor:
        self._gets += 1
        if (self._gets & 1023) == 0:
            self._purge()
        key = (w.data_ptr(), qkv)
        ent = self._store.get(key)
        storage = w.untyped_storage()
        tag = (w._version, storage._cdata, tuple(w.shape))
        if ent is not None and ent[1] == tag:
            return ent[0]
        w2 = w.detach().reshape(w.shape[0], -1)
        assert w2.dtype == torch.float32 and w2.is_contiguous()
        rows, cols = w2.shape
        sh = ent[0] if (ent is not None and ent[0].shape[0] == rows and ent[0].shape[1] == round_up(cols, 8)
                        and ent[0].device == w.device) else \
            torch.empty((rows, round_up(cols, 8)), dtype=torch.bfloat16, device=w.device)
        lib = _lib.load()
        if qkv:
            assert rows % 3 == 0
            check(lib.cream_shadow_qkv(_p(w2), _p(sh), rows // 3, cols, w2.stride(0), sh.stride(0), _stream()),
                  "cream_shadow_qkv")
        else:
            check(lib.cream_shadow_cast(_p(w2), _p(sh), rows, cols, w2.stride(0), sh.stride(0), _stream()),
                  "cream_shadow_cast")
        if ent is not None and ent[2]._cdata == storage._cdata and ent[3] is not None and ent[3]() is not None:
            wref = ent[3]                       # same parameter, new version: keep watching the original object
        else:
            try:
                wref = weakref.ref(w)
            except TypeError:
                wref = None
        self._store[key] = (sh, tag, storage, wref)
        return sh

    def clear(self):
        self._store.clear()


SHADOWS = ShadowCache()


# --------------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------------
def gemm(M, N, K, a, lda, b, ldb, out, ldo, epi, *, groups=1, a_mn=0, b_mn=0, a_group_off=0,
         b_group_rows=0, k_groups=1, k_group_len=0, out_row_mul=1, out_g_row=0, out_g_col=0, aux=None,
         ldaux=0, bias=None, resid=None, ldr=0, row_scale=None, rows_per_scale=1, alpha=1.0, split_k=0, cta_pair=0):
    d = GemmDesc()
    d.M, d.N, d.K, d.groups = M, N, K, groups
    d.a, d.lda, d.a_mn, d.a_group_off = _p(a), lda, a_mn, a_group_off
    d.b, d.ldb, d.b_mn, d.b_group_rows = _p(b), ldb, b_mn, b_group_rows
    d.k_groups, d.k_group_len = k_groups, k_group_len
    d.epi = epi
    d.out, d.ldo = _p(out), ldo
    d.out_row_mul, d.out_g_row, d.out_g_col = out_row_mul, out_g_row, out_g_col
    d.aux, d.ldaux = _p(aux), ldaux
    d.bias = _p(bias)
    d.resid, d.ldr = _p(resid), ldr
    d.row_scale, d.rows_per_scale = _p(row_scale), rows_per_scale
    d.alpha, d.split_k, d.cta_pair = alpha, split_k, cta_pair
    _profiled("gemm", 2.0 * M * N * K * groups, 0.0,
              lambda: check(_lib.load().cream_gemm_bf16(C.byref(d), _stream()), "cream_gemm_bf16"))


def linear_fwd(x, w_sh, n_out, k_in, bias=None, *, epi=EPI_BF16, out=None, aux=None, resid=None,
               row_scale=None, rows_per_scale=1):
    """y = x[:, :k_in] @ W[:n_out, :k_in]^T (+ bias[:n_out]) with the chosen epilogue
    (Linear_super.py:52-54 on the sampled slice; the slice is only extents + pitches)."""
    _check_2d(x, torch.bfloat16, "x", 8)
    M = x.shape[0]
    if out is None:
        out = empty_bf16(M, n_out, x.device) if epi in (EPI_BF16, EPI_BF16_GELU) else empty_f32(M, n_out, x.device)
    gemm(M, n_out, k_in, x, x.stride(0), w_sh, w_sh.stride(0), out, out.stride(0), epi,
         aux=aux, ldaux=aux.stride(0) if aux is not None else 0, bias=bias,
         resid=resid, ldr=resid.stride(0) if resid is not None else 0,
         row_scale=row_scale, rows_per_scale=rows_per_scale)
    return out


def linear_dgrad(dy, w_sh, n_out, k_in, *, epi=EPI_BF16, aux=None, out=None):
    """dx = dy[:, :n_out] @ W[:n_out, :k_in]  (B operand read MN-major straight from the shadow)."""
    _check_2d(dy, torch.bfloat16, "dy", 8)
    M = dy.shape[0]
    if out is None:
        out = empty_bf16(M, k_in, dy.device)
    gemm(M, k_in, n_out, dy, dy.stride(0), w_sh, w_sh.stride(0), out, out.stride(0), epi, b_mn=1,
         aux=aux, ldaux=aux.stride(0) if aux is not None else 0)
    return out


def linear_wgrad(dy, x, n_out, k_in, dw_full, alpha=1.0):
    """dW[:n_out, :k_in] += dy^T x, accumulated in place in the full-size fp32 gradient."""
    _check_2d(dy, torch.bfloat16, "dy", 8)
    _check_2d(x, torch.bfloat16, "x", 8)
    dw2 = dw_full.reshape(dw_full.shape[0], -1)
    assert dw2.dtype == torch.float32 and dw2.is_contiguous()
    gemm(n_out, k_in, dy.shape[0], dy, dy.stride(0), x, x.stride(0), dw2, dw2.stride(0), EPI_F32_ATOMIC,
         a_mn=1, b_mn=1, alpha=alpha)


def qkv_fwd(x, wq_sh, heads, k_in, rows_per_group, bias=None):
    """qkv_super.forward (qkv_super.py:45-55): three row blocks of the de-interleaved shadow."""
    _check_2d(x, torch.bfloat16, "x", 8)
    qd = HEAD_DIM * heads
    out = empty_bf16(x.shape[0], 3 * qd, x.device)
    gemm(x.shape[0], qd, k_in, x, x.stride(0), wq_sh, wq_sh.stride(0), out, out.stride(0), EPI_BF16,
         groups=3, b_group_rows=rows_per_group, out_g_col=qd, bias=bias)
    return out


def qkv_dgrad(dqkv, wq_sh, heads, k_in, rows_per_group):
    _check_2d(dqkv, torch.bfloat16, "dqkv", 8)
    qd = HEAD_DIM * heads
    out = empty_bf16(dqkv.shape[0], k_in, dqkv.device)
    gemm(dqkv.shape[0], k_in, 3 * qd, dqkv, dqkv.stride(0), wq_sh, wq_sh.stride(0), out, out.stride(0),
         EPI_BF16, b_mn=1, k_groups=3, k_group_len=qd, b_group_rows=rows_per_group)
    return out


def qkv_wgrad(dqkv, x, heads, k_in, dw_full):
    """dW[3j+i, :k_in] += dqkv[:, i*64h + j]^T x — written in the reference's interleaved rows."""
    qd = HEAD_DIM * heads
    assert dw_full.dtype == torch.float32 and dw_full.is_contiguous()
    gemm(qd, k_in, dqkv.shape[0], dqkv, dqkv.stride(0), x, x.stride(0), dw_full, dw_full.stride(0),
         EPI_F32_ATOMIC, groups=3, a_mn=1, b_mn=1, a_group_off=qd, out_row_mul=3, out_g_row=1)


def bias_grad(dy, dbias_full):
    _check_2d(dy, torch.bfloat16, "dy", 2)
    _profiled("bias_grad", 0.0, 2.0 * dy.shape[0] * dy.shape[1],
              lambda: check(_lib.load().cream_bias_grad(_p(dy), dy.stride(0), _p(dbias_full), dy.shape[0], dy.shape[1],
                                                        _stream()), "cream_bias_grad"))


def cast_scale(g, row_scale=None, rows_per_scale=1, dbias=None):
    """bf16(row_scale * g) with optional fused bias gradient (column sums)."""
    _check_2d(g, torch.float32, "g", 4)
    rows, cols = g.shape
    assert cols % 4 == 0
    out = empty_bf16(rows, cols, g.device)
    _profiled("cast_scale", 0.0, 6.0 * rows * cols,
              lambda: check(_lib.load().cream_cast_scale(_p(g), g.stride(0), _p(out), out.stride(0), _p(row_scale),
                                                         rows_per_scale, _p(dbias), rows, cols, _stream()), "cream_cast_scale"))
    return out


# --------------------------------------------------------------------------------------------
# LayerNorm
# --------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, E, *, out_f32=False, save_stats=True):
    _check_2d(x, torch.float32, "x", 1)
    rows = x.shape[0]
    out = empty_f32(rows, E, x.device) if out_f32 else empty_bf16(rows, E, x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    _profiled("ln_fwd", 0.0, rows * E * (4.0 + (4.0 if out_f32 else 2.0)),
              lambda: check(_lib.load().cream_layernorm_fwd(_p(x), x.stride(0), _p(gamma), _p(beta), eps, _p(out),
                                                            out.stride(0), int(out_f32), _p(mean), _p(rstd), rows, E,
                                                            _stream()), "cream_layernorm_fwd"))
    return out, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, E, dgamma_full, dbeta_full, resid_grad=None):
    dy_f32 = dy.dtype == torch.float32
    rows = x.shape[0]
    dx = empty_f32(rows, E, x.device)
    nbytes = rows * E * ((4.0 if dy_f32 else 2.0) + 4.0 + 4.0 + (4.0 if resid_grad is not None else 0.0))
    _profiled("ln_bwd", 0.0, nbytes,
              lambda: check(_lib.load().cream_layernorm_bwd(_p(dy), dy.stride(0), int(dy_f32), _p(x), x.stride(0), _p(gamma),
                                                            _p(mean), _p(rstd), _p(resid_grad),
                                                            resid_grad.stride(0) if resid_grad is not None else 0, _p(dx),
                                                            dx.stride(0), _p(dgamma_full), _p(dbeta_full), rows, E,
                                                            _stream()), "cream_layernorm_bwd"))
    return dx


def layernorm_bwd_cast(dy, x, gamma, mean, rstd, E, dgamma_full, dbeta_full, resid_grad=None, row_scale=None,
                       rows_per_scale=1, dbias=None):
    """layernorm_bwd that also returns bf16(row_scale * dx) and accumulates its column sums into `dbias`
    (cream_layernorm_bwd_cast): one pass instead of layernorm_bwd + cast_scale.  Returns (dx fp32, dx bf16)."""
    dy_f32 = dy.dtype == torch.float32
    rows = x.shape[0]
    dx = empty_f32(rows, E, x.device)
    out = empty_bf16(rows, E, x.device)
    nbytes = rows * E * ((4.0 if dy_f32 else 2.0) + 4.0 + 4.0 + 2.0 + (4.0 if resid_grad is not None else 0.0))
    _profiled("ln_bwd", 0.0, nbytes,
              lambda: check(_lib.load().cream_layernorm_bwd_cast(
                  _p(dy), dy.stride(0), int(dy_f32), _p(x), x.stride(0), _p(gamma), _p(mean), _p(rstd), _p(resid_grad),
                  resid_grad.stride(0) if resid_grad is not None else 0, _p(dx), dx.stride(0), _p(dgamma_full), _p(dbeta_full),
                  rows, E, _p(out), out.stride(0), _p(row_scale), rows_per_scale, _p(dbias), _stream()),
                  "cream_layernorm_bwd_cast"))
    return dx, out


# --------------------------------------------------------------------------------------------
# relative-position tables
# --------------------------------------------------------------------------------------------
_INDEX_CACHE = {}


def _u8_table(idx: np.ndarray, offset: int, device) -> torch.Tensor:
    n = idx.shape[0]
    ld = round_up(n, 16)
    buf = np.zeros((n, ld), dtype=np.uint8)
    buf[:, :n] = (idx + offset).astype(np.uint8)
    return torch.from_numpy(buf).to(device)


def autoformer_index_tables(n_tokens: int, max_rel: int, device):
    """uint8 gather tables for the AutoFormer 2-D relative position (multihead_super.py:40-59):
    idx_v in packed rows [0,32), idx_h offset into [32,64).  Built once per (N, device)."""
    key = ("af", n_tokens, max_rel, str(device))
    if key not in _INDEX_CACHE:
        grid = int(round((n_tokens - 1) ** 0.5))
        assert grid * grid + 1 == n_tokens, "AutoFormer relative position needs a square grid + cls"
        assert 2 * max_rel + 2 <= 32, "table does not fit a 32-row half of the pack"
        iv = np.empty((n_tokens, n_tokens), np.int32)
        ih = np.empty((n_tokens, n_tokens), np.int32)
        check(_lib.load().cream_autoformer_rel_index_host(grid, max_rel, iv.ctypes.data, ih.ctypes.data),
              "cream_autoformer_rel_index_host", kernels=0)
        _INDEX_CACHE[key] = (_u8_table(iv, 0, device), _u8_table(ih, 32, device), iv, ih)
    return _INDEX_CACHE[key]


def irpe_bucket_ids(method: int, height: int, width: int, skip: int, alpha: float, beta: float, gamma: float):
    """int32 bucket ids + bucket count (irpe.py:291-415), host side, cached."""
    key = ("irpe", method, height, width, skip, alpha, beta, gamma)
    if key not in _INDEX_CACHE:
        n = skip + height * width
        ids = np.empty((n, n), np.int32)
        nb = C.c_int(0)
        check(_lib.load().cream_irpe_bucket_ids_host(method, height, width, skip, alpha, beta, gamma,
                                                     ids.ctypes.data, C.byref(nb)), "cream_irpe_bucket_ids_host", kernels=0)
        _INDEX_CACHE[key] = (ids, nb.value)
    return _INDEX_CACHE[key]


def irpe_index_table_u8(ids: np.ndarray, device, offset: int = 0) -> torch.Tensor:
    """uint8 gather table of iRPE bucket ids, shifted by `offset` rows into the 64-row table pack
    (the cross method keeps its column table at rows [32, 64))."""
    key = ("irpe_u8", ids.tobytes(), offset, str(device))
    if key not in _INDEX_CACHE:
        assert ids.max() + offset < NB_PACK, "more than 64 buckets is not supported by the fused kernel"
        _INDEX_CACHE[key] = _u8_table(ids, offset, device)
    return _INDEX_CACHE[key]


def irpe_grid_product_structure(ids: np.ndarray, grid: int, skip: int):
    """If the bucket ids of a (skip + grid*grid)-token sequence have the iRPE product structure
    (irpe.py:176-202)  id(i, j) = A[rj - ri] * W + B[cj - ci]  for patch tokens and one skip bucket for
    every pair that involves the cls token, return (W, skip_id, lut_a, lut_b) with the two components
    indexed by (delta + grid - 1); otherwise None.  VERIFIED against the full table, so the kernels'
    register-arithmetic gather is bit-for-bit the table gather."""
    key = ("gp", ids.tobytes(), grid, skip)
    if key in _INDEX_CACHE:
        return _INDEX_CACHE[key]
    out = None
    n = skip + grid * grid
    if skip == 1 and ids.shape == (n, n) and grid <= 16:
        skip_id = int(ids[0, 0])
        nb = int(ids.max()) + 1
        W = int(round((nb - 1) ** 0.5))
        patch = ids[1:, 1:].reshape(grid, grid, grid, grid)                 # (ri, ci, rj, cj)
        if W * W == nb - 1 and skip_id == nb - 1 and (ids[0, :] == skip_id).all() and (ids[:, 0] == skip_id).all():
            lut_a = np.zeros(32, np.uint8)
            lut_b = np.zeros(32, np.uint8)
            for dlt in range(-(grid - 1), grid):
                r0, r1 = max(0, -dlt), max(0, dlt)
                lut_a[dlt + grid - 1] = patch[r0, 0, r1, 0] // W
                lut_b[dlt + grid - 1] = patch[0, r0, 0, r1] % W
            r = np.arange(grid)
            da = lut_a[(r[None, :] - r[:, None]) + grid - 1].astype(np.int64)      # (ri, rj)
            db = lut_b[(r[None, :] - r[:, None]) + grid - 1].astype(np.int64)      # (ci, cj)
            want = da[:, None, :, None] * W + db[None, :, None, :]
            if (want == patch).all():
                out = (W, skip_id, lut_a, lut_b)
    _INDEX_CACHE[key] = out
    return out


def pack_tables(dst, num_tables, src0, nb0, off0, strides0, src1=None, nb1=0, off1=0, strides1=(0, 0, 0)):
    check(_lib.load().cream_pack_tables(_p(dst), num_tables, HEAD_DIM, _p(src0), nb0, off0, *strides0,
                                        _p(src1), nb1, off1, *strides1, _stream()), "cream_pack_tables")


def unpack_table_grads(dpack, num_tables, g0, nb0, off0, strides0, g1=None, nb1=0, off1=0, strides1=(0, 0, 0)):
    check(_lib.load().cream_unpack_table_grads(_p(dpack), num_tables, HEAD_DIM, _p(g0), nb0, off0, *strides0,
                                               _p(g1), nb1, off1, *strides1, _stream()),
          "cream_unpack_table_grads")


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def pack_tables_batch(pairs, device) -> torch.Tensor:
    """pairs: list of (table_a, table_b) fp32 (nb, 64) tensors -> (len, 64, 64) bf16 packs, one launch."""
    t0 = pairs[0][0]
    nb = t0.shape[0]
    dst = torch.empty((len(pairs), NB_PACK, HEAD_DIM), dtype=torch.bfloat16, device=device)
    a0, a1 = _ptr_array([p[0] for p in pairs]), _ptr_array([p[1] for p in pairs])
    check(_lib.load().cream_pack_tables_batch(_p(dst), len(pairs), HEAD_DIM, a0, a1, nb, 32, t0.stride(0), t0.stride(1),
                                              _stream()), "cream_pack_tables_batch")
    return dst


def unpack_table_grads_batch(dpacks: torch.Tensor, grad_pairs) -> None:
    g0 = grad_pairs[0][0]
    a0, a1 = _ptr_array([p[0] for p in grad_pairs]), _ptr_array([p[1] for p in grad_pairs])
    check(_lib.load().cream_unpack_table_grads_batch(_p(dpacks), len(grad_pairs), HEAD_DIM, a0, a1, g0.shape[0], 32,
                                                     g0.stride(0), g0.stride(1), _stream()),
          "cream_unpack_table_grads_batch")


def new_pack(num_tables: int, device) -> torch.Tensor:
    return torch.empty((num_tables, NB_PACK, HEAD_DIM), dtype=torch.bfloat16, device=device)


# --------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------
def _set_dense(d, dense, B, H, N):
    """dense: fp32 (B|1, H|1, N, N) additive logit term, last dim contiguous; size-1 dims broadcast."""
    if dense is None:
        return
    if dense.dtype != torch.float32 or not dense.is_cuda or dense.dim() != 4 or dense.stride(3) != 1:
        raise RuntimeError("cream_b200: dense attention bias must be a CUDA fp32 (B|1, H|1, N, N) tensor, last dim contiguous")
    assert dense.shape[2] == N and dense.shape[3] == N and dense.shape[0] in (1, B) and dense.shape[1] in (1, H)
    d.dense_bias = _p(dense)
    d.dense_stride_b = dense.stride(0) if dense.shape[0] > 1 else 0
    d.dense_stride_h = dense.stride(1) if dense.shape[1] > 1 else 0
    d.dense_stride_i = dense.stride(2)


def _attn_desc(B, H, N, scale, qkv, tk, tv, per_head, idx, bias, af=None, gp=None, causal=False, block=0):
    d = AttnDesc()
    d.causal = int(bool(causal))
    d.block_len = int(block)
    assert block == 0 or (N % block == 0 and af is None and gp is None), "block-diagonal attention: N % block_len == 0, generic path"
    if af is not None:
        d.af_grid, d.af_max_rel = af
    if gp is not None:          # (grid, W, skip_id, lut_a, lut_b) from irpe_grid_product_structure
        d.gp_grid, d.gp_w, d.gp_skip_id = gp[0], gp[1], gp[2]
        C.memmove(d.gp_lut_a, gp[3].ctypes.data, 32)
        C.memmove(d.gp_lut_b, gp[4].ctypes.data, 32)
    d.B, d.H, d.N, d.head_dim = B, H, N, HEAD_DIM
    d.scale = scale
    d.qkv, d.ld_qkv = _p(qkv), qkv.stride(0)
    d.tk_pack, d.tv_pack, d.tables_per_head = _p(tk), _p(tv), int(per_head)
    ia, ib, iva, ivb = idx
    d.idx_a, d.idx_b, d.idx_va, d.idx_vb = _p(ia), _p(ib), _p(iva), _p(ivb)
    first = next((t for t in idx if t is not None), None)
    d.ld_idx = first.stride(0) if first is not None else 0
    d.bias_pack = _p(bias)
    return d


def attention_fwd(qkv, B, H, N, scale, *, tk=None, tv=None, per_head=False, idx=(None, None, None, None),
                  bias=None, need_lse=True, af=None, dense=None, gp=None, causal=False, block=0):
    """Fused attention forward.  qkv: (B*N, 3*H*64) bf16.  Returns (out (B*N, H*64) bf16, lse).
    dense: optional fp32 (B|1, H|1, N, N) term added to the logits (see cream_attn_desc.dense_bias).
    block: > 0 = block-diagonal attention over items of `block` tokens (see cream_attn_desc.block_len)."""
    _check_2d(qkv, torch.bfloat16, "qkv", 8)
    out = empty_bf16(B * N, H * HEAD_DIM, qkv.device)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device) if need_lse else None
    d = _attn_desc(B, H, N, scale, qkv, tk, tv, per_head, idx, bias, af, gp, causal, block)
    _set_dense(d, dense, B, H, N)
    d.out, d.ld_out, d.lse = _p(out), out.stride(0), _p(lse)
    nb = (NB_PACK if tk is not None else 0) + (NB_PACK if tv is not None else 0)
    _profiled("attn_fwd", 4.0 * B * H * N * N * HEAD_DIM + 2.0 * B * H * N * HEAD_DIM * nb, 4.0 * B * H * N * HEAD_DIM * 2,
              lambda: check(_lib.load().cream_attn_fwd(C.byref(d), _stream()), "cream_attn_fwd"))
    return out, lse


def attention_bwd(qkv, out, lse, dout, B, H, N, scale, *, tk=None, tv=None, per_head=False,
                  idx=(None, None, None, None), bias=None, af=None, dtk=None, dtv=None, dense=None, ddense=None, gp=None,
                  causal=False, block=0):
    """Returns (dqkv bf16, dtk_pack fp32|None, dtv_pack fp32|None, dbias fp32|None).  dtk / dtv may
    be caller-provided zeroed (T, 64, 64) fp32 accumulators.  With `dense`, pass `ddense` = an fp32
    (B, H, N, N) tensor to receive the logit gradient dS."""
    _check_2d(dout, torch.bfloat16, "dout", 8)
    dev = qkv.device
    dqkv = empty_bf16(B * N, 3 * H * HEAD_DIM, dev)
    T = H if per_head else 1
    if dtk is None and tk is not None:
        dtk = torch.zeros((T, NB_PACK, HEAD_DIM), dtype=torch.float32, device=dev)
    if dtv is None and tv is not None:
        dtv = torch.zeros((T, NB_PACK, HEAD_DIM), dtype=torch.float32, device=dev)
    dbias = torch.zeros((T, NB_PACK), dtype=torch.float32, device=dev) if bias is not None else None
    nbytes = _lib.load().cream_attn_bwd_workspace_bytes(B, H, N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d = _attn_desc(B, H, N, scale, qkv, tk, tv, per_head, idx, bias, af, gp, causal, block)
    d.out, d.ld_out, d.lse = _p(out), out.stride(0), _p(lse)
    d.dout, d.ld_dout = _p(dout), dout.stride(0)
    d.dqkv, d.ld_dqkv = _p(dqkv), dqkv.stride(0)
    d.dtk_pack, d.dtv_pack, d.dbias_pack = _p(dtk), _p(dtv), _p(dbias)
    d.workspace, d.workspace_bytes = _p(ws), nbytes
    _set_dense(d, dense, B, H, N)
    if ddense is not None:
        assert ddense.dtype == torch.float32 and ddense.is_contiguous() and tuple(ddense.shape) == (B, H, N, N)
        d.ddense = _p(ddense)
    nb = (NB_PACK if tk is not None else 0) + (NB_PACK if tv is not None else 0)
    # backward = 2.5 x forward FLOPs (S recomputed); algorithmic bytes: q, k, v, o, do in, dq, dk, dv out
    _profiled("attn_bwd", 2.5 * (4.0 * B * H * N * N * HEAD_DIM + 2.0 * B * H * N * HEAD_DIM * nb),
              8.0 * B * H * N * HEAD_DIM * 2,
              lambda: check(_lib.load().cream_attn_bwd(C.byref(d), _stream()), "cream_attn_bwd", kernels=2))
    return dqkv, dtk, dtv, dbias


# --------------------------------------------------------------------------------------------
# rpe_index (the reference's native op)
# --------------------------------------------------------------------------------------------
_DT = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16,
       torch.float64: _lib.DTYPE_F64}


def rpe_index_forward(inp: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """rpe_index_cpp.forward_gpu (rpe_index_cuda.cu:54-94): same checks, same semantics."""
    if not inp.is_cuda:
        raise RuntimeError("input must be a GPU tensor")
    if not index.is_cuda:
        raise RuntimeError("index must be a GPU tensor")
    if inp.dim() != 4:
        raise RuntimeError("input must be a 4D tensor")
    if index.dim() != 2:
        raise RuntimeError("index must be a 2D tensor")
    if index.dtype != torch.int32:
        raise RuntimeError("index must be Int type")
    if not index.is_contiguous():
        raise RuntimeError("index should be contiguous")
    if inp.dtype not in _DT:
        raise RuntimeError(f"rpe_index: unsupported dtype {inp.dtype}")
    B, H, _, nb = inp.shape
    Lq, Lk = index.shape
    out = torch.empty((B, H, Lq, Lk), dtype=inp.dtype, device=inp.device)
    s = inp.stride()
    with torch.cuda.device(inp.device):
        check(_lib.load().cream_rpe_index_fwd(_p(inp), _p(index), _p(out), B, H, Lq, Lk, nb, s[0], s[1], s[2], s[3],
                                              _DT[inp.dtype], _stream()), "cream_rpe_index_fwd")
    return out


def rpe_index_backward(grad_input: torch.Tensor, grad_output: torch.Tensor, index: torch.Tensor) -> None:
    """rpe_index_cpp.backward_gpu (rpe_index_cuda.cu:96-140): accumulates into grad_input."""
    if not (grad_input.is_cuda and grad_output.is_cuda and index.is_cuda):
        raise RuntimeError("grad_input, grad_output and index must be GPU tensors")
    if grad_input.dim() != 4 or grad_output.dim() != 4 or index.dim() != 2:
        raise RuntimeError("input must be a 4D tensor / index must be a 2D tensor")
    if index.dtype != torch.int32:
        raise RuntimeError("index must be Int type")
    gi = grad_input if grad_input.is_contiguous() else grad_input.contiguous()
    go = grad_output.contiguous()
    idx = index.contiguous()
    B, H, Lq, Lk = go.shape
    nb = gi.shape[3]
    with torch.cuda.device(go.device):
        check(_lib.load().cream_rpe_index_bwd(_p(gi), _p(go), _p(idx), B, H, Lq, Lk, nb, _DT[go.dtype], _stream()),
              "cream_rpe_index_bwd")
    if gi is not grad_input:
        grad_input.copy_(gi)
