"""torch.autograd.Functions behind the drop-in modules (AutoFormer model.module.* and the
iRPE attention).  Each forward/backward is a handful of cream_b200 C-ABI launches; torch
only carries tensors and the autograd graph.  Gradients of sliced parameters are returned
as full-size tensors that are zero outside the sampled slice, exactly like autograd through
the reference's slicing views."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, ops
from .._lib import check

_p, _stream = ops._p, ops._stream


def _f32_2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != torch.float32 or x2.stride(1) != 1:
        x2 = x2.float().contiguous()
    return x2


class SlicedLinearFn(torch.autograd.Function):
    """y = x @ W[:out, :in]^T + b[:out]  (Linear_super.py:52-54) or, with is_qkv, the
    interleaved-row QKV slice of qkv_super.py:45-55,72-83."""

    @staticmethod
    def forward(ctx, x, weight, bias, in_dim, out_dim, is_qkv):
        if not x.is_cuda:
            raise RuntimeError("cream_b200 modules run on CUDA (sm_100a) tensors only")
        assert x.shape[-1] == in_dim, f"input feature dim {x.shape[-1]} != sampled in_dim {in_dim}"
        x2 = ops.as_bf16_2d(x)
        if is_qkv:
            assert out_dim % (3 * ops.HEAD_DIM) == 0, "sampled qkv width must be 3*64*heads"
            heads = out_dim // (3 * ops.HEAD_DIM)
            y = ops.qkv_fwd(x2, ops.SHADOWS.get(weight, qkv=True), heads, in_dim, weight.shape[0] // 3, bias)
        else:
            y = ops.linear_fwd(x2, ops.SHADOWS.get(weight), out_dim, in_dim, bias)
        ctx.save_for_backward(x2, weight)
        ctx.meta = (in_dim, out_dim, is_qkv, bias is not None, x.dtype, tuple(x.shape))
        return y.reshape(*x.shape[:-1], out_dim)

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        in_dim, out_dim, is_qkv, has_bias, x_dtype, x_shape = ctx.meta
        dy2 = ops.as_bf16_2d(dy)
        dx = dw = db = None
        if is_qkv:
            heads = out_dim // (3 * ops.HEAD_DIM)
            if ctx.needs_input_grad[0]:
                dx = ops.qkv_dgrad(dy2, ops.SHADOWS.get(weight, qkv=True), heads, in_dim, weight.shape[0] // 3)
            if ctx.needs_input_grad[1]:
                dw = torch.zeros_like(weight)
                ops.qkv_wgrad(dy2, x2, heads, in_dim, dw)
        else:
            if ctx.needs_input_grad[0]:
                dx = ops.linear_dgrad(dy2, ops.SHADOWS.get(weight), out_dim, in_dim)
            if ctx.needs_input_grad[1]:
                dw = torch.zeros_like(weight)
                ops.linear_wgrad(dy2, x2, out_dim, in_dim, dw)
        if has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(weight.shape[0], dtype=torch.float32, device=weight.device)
            ops.bias_grad(dy2, db)
        if dx is not None:
            dx = dx.reshape(x_shape).to(x_dtype)
        return dx, dw, db, None, None, None


class SlicedLayerNormFn(torch.autograd.Function):
    """F.layer_norm over the first E features with weight[:E], bias[:E]
    (layernorm_super.py:35-37); fp32 in, fp32 out (as under the reference's autocast)."""

    @staticmethod
    def forward(ctx, x, weight, bias, E, eps):
        if not x.is_cuda:
            raise RuntimeError("cream_b200 modules run on CUDA (sm_100a) tensors only")
        assert x.shape[-1] == E
        x2 = _f32_2d(x)
        y, mean, rstd = ops.layernorm_fwd(x2, weight, bias, eps, E, out_f32=True)
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.meta = (E, x.dtype, tuple(x.shape))
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        E, x_dtype, x_shape = ctx.meta
        dy2 = _f32_2d(dy)
        dg = torch.zeros_like(weight)
        db = torch.zeros_like(weight)
        dx = ops.layernorm_bwd(dy2, x2, weight, mean, rstd, E, dg, db)
        return dx.reshape(x_shape).to(x_dtype), dg, db, None, None


class PatchEmbedFn(torch.autograd.Function):
    """conv2d(x, W[:E], b[:E], stride=P).flatten(2).transpose(1, 2) (embedding_super.py:33-40)
    as im2col + sliced GEMM; returns (B, T, E) bf16."""

    @staticmethod
    def forward(ctx, images, weight, bias, E, P):
        if not images.is_cuda:
            raise RuntimeError("cream_b200 modules run on CUDA (sm_100a) tensors only")
        B, Cin, H, W = images.shape
        img = images.float().contiguous()
        T = (H // P) * (W // P)
        kdim = Cin * P * P
        cols = ops.empty_bf16(B * T, kdim, images.device)
        check(_lib.load().cream_patch_im2col(_p(img), _p(cols), cols.stride(0), B, Cin, H, W, P, _stream()),
              "cream_patch_im2col")
        y = ops.linear_fwd(cols, ops.SHADOWS.get(weight), E, kdim, bias)
        ctx.save_for_backward(cols, weight)
        ctx.meta = (E, kdim, bias is not None)
        return y.reshape(B, T, E)

    @staticmethod
    def backward(ctx, dy):
        cols, weight = ctx.saved_tensors
        E, kdim, has_bias = ctx.meta
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("gradient w.r.t. the input image is not part of the supernet path")
        dy2 = ops.as_bf16_2d(dy)
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw = torch.zeros_like(weight)
            ops.linear_wgrad(dy2, cols, E, kdim, dw)
        if has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(weight.shape[0], dtype=torch.float32, device=weight.device)
            ops.bias_grad(dy2, db)
        return None, dw, db, None, None


def _pack_pair(tv, th):
    pack = ops.new_pack(1, tv.device)
    nb = tv.shape[0]
    ops.pack_tables(pack, 1, tv, nb, 0, (0, tv.stride(0), tv.stride(1)), th, nb, 32, (0, th.stride(0), th.stride(1)))
    return pack


class AutoformerAttentionFn(torch.autograd.Function):
    """Attention core of AttentionSuper.forward (multihead_super.py:135-154) on a
    (B, N, 3*64h) qkv tensor; optional tables = (k_v, k_h, v_v, v_h) each (2*max_rel+2, 64)."""

    @staticmethod
    def forward(ctx, qkv, heads, scale, max_rel, *tables):
        B, N, W3 = qkv.shape
        assert W3 == 3 * ops.HEAD_DIM * heads
        qkv2 = ops.as_bf16_2d(qkv)
        tk = tv = af = None
        idx = (None, None, None, None)
        if tables:
            assert tables[0].shape[1] == ops.HEAD_DIM and tables[0].dtype == torch.float32
            iv, ih, _, _ = ops.autoformer_index_tables(N, max_rel, qkv.device)
            idx = (iv, ih, iv, ih)
            af = (int(round((N - 1) ** 0.5)), max_rel)
            tk = _pack_pair(tables[0].detach(), tables[1].detach())
            tv = _pack_pair(tables[2].detach(), tables[3].detach())
        out, lse = ops.attention_fwd(qkv2, B, heads, N, scale, tk=tk, tv=tv, idx=idx, af=af)
        ctx.save_for_backward(qkv2, out, lse, tk, tv, *tables)
        ctx.meta = (B, heads, N, scale, idx, qkv.dtype, af)
        return out.reshape(B, N, ops.HEAD_DIM * heads)

    @staticmethod
    def backward(ctx, dout):
        qkv2, out, lse, tk, tv, *tables = ctx.saved_tensors
        B, heads, N, scale, idx, dtype, af = ctx.meta
        d2 = ops.as_bf16_2d(dout)
        dqkv, dtk, dtv, _ = ops.attention_bwd(qkv2, out, lse, d2, B, heads, N, scale, tk=tk, tv=tv, idx=idx, af=af)
        grads = []
        if tables:
            for pair, dpack in ((tables[0:2], dtk), (tables[2:4], dtv)):
                gv, gh = torch.zeros_like(pair[0]), torch.zeros_like(pair[1])
                ops.unpack_table_grads(dpack, 1, gv, gv.shape[0], 0, (0, gv.stride(0), gv.stride(1)),
                                       gh, gh.shape[0], 32, (0, gh.stride(0), gh.stride(1)))
                grads += [gv, gh]
        return (dqkv.reshape(B, N, -1).to(dtype), None, None, None, *grads)


class IrpeAttentionFn(torch.autograd.Function):
    """Attention core of RPEAttention.forward (rpe_vision_transformer.py:73-92) with iRPE on
    keys and/or values.  qkv (B, N, 3*64h); rpe_k: contextual lookup_table_weight
    (H|1, 64, nb) or bias lookup_table_bias (H|1, nb); rpe_v: (H|1, nb, 64); ids: int (N, N)
    bucket ids (numpy, from ops.irpe_bucket_ids).
    Cross method (iRPE_Cross, irpe.py:696-751): ids = (row ids, column ids) and rpe_k2 / rpe_v2 are
    the column tables; the two tables share one 64-row pack (rows at [0,32), columns at [32,64))
    and the kernel gathers from both, exactly like AutoFormer's vertical / horizontal pair."""

    @staticmethod
    def _q_term(qkv2, B, N, heads, scale, ids_list, tables, mode):
        """rpe_q(k * scale)^T of RPEAttention.forward (rpe_vision_transformer.py:82-83) as a dense
        (B|1, H, N, N) logit term: lookup = (k*scale) @ table (irpe.py:617-633), gathered per KEY row
        with the library's own rpe_index kernel, transposed to (query, key).  Returns (dense, saved)."""
        dev = qkv2.device
        if mode == "bias":
            dense = None
            for ids, tab in zip(ids_list, tables):
                idl = torch.from_numpy(ids.astype(np.int64)).to(dev)
                t = tab[:, idl.flatten()].view(1, tab.shape[0], N, N).transpose(2, 3)       # (1, T, i, j)
                dense = t if dense is None else dense + t
            return dense.expand(1, heads, N, N).contiguous().float(), None
        k = qkv2.view(B, N, 3, heads, ops.HEAD_DIM)[:, :, 1].permute(0, 2, 1, 3).float() * scale   # (B, H, N, D)
        dense, saved = None, []
        for ids, tab in zip(ids_list, tables):
            w = tab.detach().float()                                                      # (T, D, nb)
            lookup = torch.matmul(k, w.unsqueeze(0)).contiguous()                         # (B, H, N, nb)
            idx32 = torch.from_numpy(ids.astype(np.int32)).to(dev)
            y = ops.rpe_index_forward(lookup, idx32)                                      # [b,h,j,i] = lookup[b,h,j,ids[j,i]]
            dense = y if dense is None else dense + y
            saved.append((idx32, lookup.shape))
        return dense.transpose(2, 3).contiguous(), (k, saved)

    @staticmethod
    def forward(ctx, qkv, heads, scale, ids, mode, rpe_k, rpe_v, rpe_k2=None, rpe_v2=None, rpe_q=None, rpe_q2=None, block=0):
        """block > 0: the N tokens of a sequence are N / block independent items (cream_attn_desc.block_len); `ids` is
        then the (N, N) table of the packed sequence (the caller tiles the per-item table)."""
        B, N, W3 = qkv.shape
        assert W3 == 3 * ops.HEAD_DIM * heads
        dev = qkv.device
        qkv2 = ops.as_bf16_2d(qkv)
        cross = isinstance(ids, (tuple, list))
        half = ops.NB_PACK // 2
        if cross:
            assert mode != "bias", "cross + bias mode is not supported by the fused kernel"
            idx_t = ops.irpe_index_table_u8(ids[0], dev)
            idx_t2 = ops.irpe_index_table_u8(ids[1], dev, offset=half)
            nb = int(max(ids[0].max(), ids[1].max())) + 1
            assert nb <= half, "cross tables must fit 32 packed rows each"
        else:
            idx_t, idx_t2 = ops.irpe_index_table_u8(ids, dev), None
            nb = int(ids.max()) + 1
        tk = tv = bias = None
        per_head = False
        if rpe_k is not None:
            per_head = rpe_k.shape[0] > 1
            T = rpe_k.shape[0]
            if mode == "bias":
                bias = torch.zeros((T, ops.NB_PACK), dtype=torch.float32, device=dev)
                bias[:, :rpe_k.shape[1]] = rpe_k.detach()
            else:
                assert rpe_k.shape[1] == ops.HEAD_DIM and rpe_k.shape[2] >= nb
                tk = ops.new_pack(T, dev)
                w = rpe_k.detach()   # (T, D, nb): bucket stride = stride(2), channel stride = stride(1)
                if cross:
                    w2 = rpe_k2.detach()
                    ops.pack_tables(tk, T, w, w.shape[2], 0, (w.stride(0), w.stride(2), w.stride(1)),
                                    w2, w2.shape[2], half, (w2.stride(0), w2.stride(2), w2.stride(1)))
                else:
                    ops.pack_tables(tk, T, w, w.shape[2], 0, (w.stride(0), w.stride(2), w.stride(1)))
        if rpe_v is not None:
            assert mode != "bias", "bias mode has no value-side table (irpe.py:489-491)"
            per_head = per_head or rpe_v.shape[0] > 1
            T = rpe_v.shape[0]
            assert rpe_k is None or rpe_k.shape[0] == T, "k and v tables must agree on head sharing"
            tv = ops.new_pack(T, dev)
            w = rpe_v.detach()       # (T, nb, D)
            if cross:
                w2 = rpe_v2.detach()
                ops.pack_tables(tv, T, w, w.shape[1], 0, (w.stride(0), w.stride(1), w.stride(2)),
                                w2, w2.shape[1], half, (w2.stride(0), w2.stride(1), w2.stride(2)))
            else:
                ops.pack_tables(tv, T, w, w.shape[1], 0, (w.stride(0), w.stride(1), w.stride(2)))
        on_k = tk is not None or bias is not None
        idx = (idx_t if on_k else None, idx_t2 if on_k else None,
               idx_t if tv is not None else None, idx_t2 if tv is not None else None)
        dense = qsave = None
        if rpe_q is not None:
            ids_list = list(ids) if cross else [ids]
            dense, qsave = IrpeAttentionFn._q_term(qkv2, B, N, heads, scale, ids_list,
                                                   [rpe_q, rpe_q2] if cross else [rpe_q], mode)
        gp = None
        if not cross and tk is not None and tv is None and bias is None and dense is None:
            side = int(round((N - 1) ** 0.5))
            st = ops.irpe_grid_product_structure(ids, side, N - side * side) if side * side + 1 == N else None
            if st is not None:
                gp = (side,) + st
        if block:
            gp = None
        out, lse = ops.attention_fwd(qkv2, B, heads, N, scale, tk=tk, tv=tv, per_head=per_head, idx=idx, bias=bias,
                                     dense=dense, gp=gp, block=block)
        ctx.save_for_backward(qkv2, out, lse, tk, tv, bias, rpe_k, rpe_v, rpe_k2, rpe_v2, rpe_q, rpe_q2, dense)
        ctx.meta = (B, heads, N, scale, idx, per_head, mode, qkv.dtype, cross, ids, qsave, gp, block)
        return out.reshape(B, N, ops.HEAD_DIM * heads)

    @staticmethod
    def backward(ctx, dout):
        qkv2, out, lse, tk, tv, bias, rpe_k, rpe_v, rpe_k2, rpe_v2, rpe_q, rpe_q2, dense = ctx.saved_tensors
        B, heads, N, scale, idx, per_head, mode, dtype, cross, ids, qsave, gp, block = ctx.meta
        half = ops.NB_PACK // 2
        d2 = ops.as_bf16_2d(dout)
        ddense = torch.empty((B, heads, N, N), dtype=torch.float32, device=qkv2.device) if dense is not None else None
        dqkv, dtk, dtv, dbias = ops.attention_bwd(qkv2, out, lse, d2, B, heads, N, scale, tk=tk, tv=tv,
                                                  per_head=per_head, idx=idx, bias=bias, dense=dense, ddense=ddense, gp=gp,
                                                  block=block)
        gk = gv = gk2 = gv2 = gq = gq2 = None
        if rpe_q is not None:
            tabs = [rpe_q, rpe_q2] if cross else [rpe_q]
            ids_list = list(ids) if cross else [ids]
            grads = []
            dy = ddense.transpose(2, 3).contiguous()                     # [b,h,j,i]: gradient of the gathered lookup
            if mode == "bias":
                for idn, tab in zip(ids_list, tabs):
                    g = torch.zeros_like(tab)                            # (T, nb)
                    src = dy.sum(0) if tab.shape[0] > 1 else dy.sum((0, 1))[None]
                    g.view(tab.shape[0], -1).index_add_(1, torch.from_numpy(idn.astype(np.int64)).flatten().to(dy.device),
                                                        src.reshape(tab.shape[0], -1))
                    grads.append(g)
            else:
                kq, saved = qsave
                dk = torch.zeros_like(kq)
                for (idx32, lshape), tab in zip(saved, tabs):
                    dl = torch.zeros(lshape, dtype=torch.float32, device=dy.device)
                    ops.rpe_index_backward(dl, dy, idx32)                # scatter-add per key row
                    w = tab.detach().float()                             # (T, D, nb)
                    gw = torch.einsum("bhnd,bhnc->hdc", kq, dl)
                    grads.append((gw if tab.shape[0] > 1 else gw.sum(0, keepdim=True)).to(tab.dtype))
                    dk += torch.matmul(dl, w.transpose(1, 2).unsqueeze(0))
                dq5 = dqkv.view(B, N, 3, heads, ops.HEAD_DIM)
                dq5[:, :, 1] += (dk * scale).permute(0, 2, 1, 3).to(dqkv.dtype)
            gq = grads[0]
            gq2 = grads[1] if cross else None
        if rpe_k is not None:
            gk = torch.zeros_like(rpe_k)
            if mode == "bias":
                gk += dbias[:, :rpe_k.shape[1]]
            else:
                T = rpe_k.shape[0]
                if cross:
                    gk2 = torch.zeros_like(rpe_k2)
                    ops.unpack_table_grads(dtk, T, gk, gk.shape[2], 0, (gk.stride(0), gk.stride(2), gk.stride(1)),
                                           gk2, gk2.shape[2], half, (gk2.stride(0), gk2.stride(2), gk2.stride(1)))
                else:
                    ops.unpack_table_grads(dtk, T, gk, gk.shape[2], 0, (gk.stride(0), gk.stride(2), gk.stride(1)))
        if rpe_v is not None:
            gv = torch.zeros_like(rpe_v)
            T = rpe_v.shape[0]
            if cross:
                gv2 = torch.zeros_like(rpe_v2)
                ops.unpack_table_grads(dtv, T, gv, gv.shape[1], 0, (gv.stride(0), gv.stride(1), gv.stride(2)),
                                       gv2, gv2.shape[1], half, (gv2.stride(0), gv2.stride(1), gv2.stride(2)))
            else:
                ops.unpack_table_grads(dtv, T, gv, gv.shape[1], 0, (gv.stride(0), gv.stride(1), gv.stride(2)))
        return dqkv.reshape(B, N, -1).to(dtype), None, None, None, None, gk, gv, gk2, gv2, gq, gq2, None


class DenseAttentionFn(torch.autograd.Function):
    """softmax(scale * q k^T + dense) v with the fused kernel; qkv (B, N, 3*64h) in the reference's
    [q | k | v] column order, dense an optional fp32 (B|1, H|1, N, N) additive logit term (mask or
    bias).  Returns (B, N, 64h).  The gradient of `dense` is produced only when it requires grad."""

    @staticmethod
    def forward(ctx, qkv, heads, scale, dense):
        B, N, W3 = qkv.shape
        assert W3 == 3 * ops.HEAD_DIM * heads
        qkv2 = ops.as_bf16_2d(qkv)
        out, lse = ops.attention_fwd(qkv2, B, heads, N, scale, dense=dense)
        ctx.save_for_backward(qkv2, out, lse, dense)
        ctx.meta = (B, heads, N, scale, qkv.dtype, dense is not None and dense.requires_grad)
        return out.reshape(B, N, ops.HEAD_DIM * heads)

    @staticmethod
    def backward(ctx, dout):
        qkv2, out, lse, dense = ctx.saved_tensors
        B, heads, N, scale, dtype, want = ctx.meta
        ddense = torch.empty((B, heads, N, N), dtype=torch.float32, device=qkv2.device) if want else None
        dqkv, _, _, _ = ops.attention_bwd(qkv2, out, lse, ops.as_bf16_2d(dout), B, heads, N, scale,
                                          dense=dense, ddense=ddense)
        gd = None
        if want:
            gd = ddense
            if dense.shape[0] == 1:
                gd = gd.sum(0, keepdim=True)
            if dense.shape[1] == 1:
                gd = gd.sum(1, keepdim=True)
        return dqkv.reshape(B, N, -1).to(dtype), None, None, gd
