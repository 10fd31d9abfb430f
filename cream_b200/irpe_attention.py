"""DeiT attention with image relative position encoding on the fused B200 kernel.

`RPEAttention` mirrors iRPE/DeiT-with-iRPE/rpe_vision_transformer.py:45-97 (constructor,
parameter names `qkv`, `proj`, `rpe_q/rpe_k/rpe_v.lookup_table_{weight,bias}`), with the
reference's q@k^T + rpe_k(q) gather, softmax, attn@v + rpe_v(attn) executed by ONE kernel.
Supported: rpe on k and/or v (contextual), bias mode on k, shared or per-head tables, methods
euclidean / quant / product (<= 64 buckets) and cross (rows + columns tables, contextual).
iRPE on queries (rpe_q) rides on the kernel's dense additive logit term: its lookup GEMM is a
library matmul, its gather / scatter-add the library's own rpe_index kernels.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .autoformer.functional import IrpeAttentionFn, SlicedLinearFn

METHODS = {"euc": 0, "quant": 1, "product": 3, "cross": 4}
CROSS_ROWS, CROSS_COLS = 41, 42


class IrpeTable(nn.Module):
    """Parameter holder with the reference's iRPE attribute names (irpe.py:449-496)."""

    def __init__(self, head_dim, num_heads, mode, transposed, num_buckets):
        super().__init__()
        self.head_dim, self.num_heads, self.mode = head_dim, num_heads, mode
        self.transposed, self.num_buckets = transposed, num_buckets
        if transposed:
            if mode == 'bias':
                self.lookup_table_bias = nn.Parameter(torch.zeros(num_heads, num_buckets))
            else:
                self.lookup_table_weight = nn.Parameter(torch.zeros(num_heads, head_dim, num_buckets))
        else:
            if mode == 'bias':
                raise NotImplementedError("[Error] Bias non-transposed RPE does not exist.")
            self.lookup_table_weight = nn.Parameter(torch.zeros(num_heads, num_buckets, head_dim))

    @property
    def table(self):
        return self.lookup_table_bias if self.mode == 'bias' else self.lookup_table_weight


class IrpeCrossTable(nn.Module):
    """iRPE_Cross (irpe.py:696-751): `rp_rows` + `rp_cols`, each an iRPE table of its own."""

    def __init__(self, head_dim, num_heads, mode, transposed, num_buckets):
        super().__init__()
        self.rp_rows = IrpeTable(head_dim, num_heads, mode, transposed, num_buckets)
        self.rp_cols = IrpeTable(head_dim, num_heads, mode, transposed, num_buckets)


class RPEAttention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.,
                 rpe_on='k', method='product', mode='contextual', shared_head=True, ratio=1.9, skip=1):
        super().__init__()
        assert attn_drop == 0.0, "attention dropout is not supported by the fused kernel"
        self.num_heads = num_heads
        head_dim = dim // num_heads
        assert head_dim == ops.HEAD_DIM, "fused attention kernel is built for head_dim 64"
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.method, self.mode, self.skip, self.ratio = METHODS[method], mode, skip, ratio
        if mode == 'ctx':
            self.mode = 'contextual'
        beta_int = int(2 * ratio)
        nb = ((2 * beta_int + 1) ** 2 if method == 'product' else 2 * beta_int + 1) + (1 if skip > 0 else 0)
        self.cross = method == 'cross'
        assert nb <= (ops.NB_PACK // 2 if self.cross else ops.NB_PACK), "too many buckets for the fused kernel"
        if self.cross and self.mode == 'bias':
            raise NotImplementedError("cross + bias mode is not supported by the fused kernel")
        self.num_buckets = nb
        t_heads = 1 if shared_head else num_heads
        make = IrpeCrossTable if self.cross else IrpeTable
        self.rpe_q = make(head_dim, t_heads, self.mode, True, nb) if 'q' in rpe_on else None
        self.rpe_k = make(head_dim, t_heads, self.mode, True, nb) if 'k' in rpe_on else None
        self.rpe_v = make(head_dim, t_heads, self.mode, False, nb) if 'v' in rpe_on else None

    def bucket_ids(self, L, height=None, width=None):
        """Bucket ids of an L-token sequence.  By default the grid is floor(sqrt(L)) square and the
        remaining tokens are skipped (irpe.py:550-566); `height` / `width` give a non-square grid as the
        DETR copy passes them (iRPE/DETR-with-iRPE/models/rpe_attention/rpe_attention_function.py:327-376)."""
        if height is None:
            height = width = int(math.sqrt(L))
        skip = L - height * width
        assert skip >= 0, "height * width exceeds the sequence length"
        if self.cross:
            out = []
            for m in (CROSS_ROWS, CROSS_COLS):
                ids, nb = ops.irpe_bucket_ids(m, height, width, skip, 1 * self.ratio, 2 * self.ratio, 8 * self.ratio)
                assert nb == self.num_buckets
                out.append(ids)
            return tuple(out)
        ids, nb = ops.irpe_bucket_ids(self.method, height, width, skip, 1 * self.ratio, 2 * self.ratio, 8 * self.ratio)
        assert nb == self.num_buckets
        return ids

    def forward(self, x, height=None, width=None):
        B, N, C = x.shape
        qkv = SlicedLinearFn.apply(x, self.qkv.weight, self.qkv.bias, C, 3 * C, False)   # (B, N, 3C) bf16
        pick = (lambda t: t.rp_rows.table) if self.cross else (lambda t: t.table)
        pick2 = (lambda t: t.rp_cols.table) if self.cross else (lambda t: None)
        first = [pick(t) if t is not None else None for t in (self.rpe_k, self.rpe_v)]
        second = [pick2(t) if t is not None else None for t in (self.rpe_k, self.rpe_v)]
        q1 = pick(self.rpe_q) if self.rpe_q is not None else None
        q2 = pick2(self.rpe_q) if self.rpe_q is not None else None
        out = IrpeAttentionFn.apply(qkv, self.num_heads, float(self.scale), self.bucket_ids(N, height, width), self.mode,
                                    *first, *second, q1, q2)
        out = SlicedLinearFn.apply(out, self.proj.weight, self.proj.bias, C, C, False)
        return self.proj_drop(out)
