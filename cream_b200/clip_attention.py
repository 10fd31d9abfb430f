"""TinyCLIP transformer-block attention on the fused B200 kernel.

`ClipAttention` mirrors the attention of TinyCLIP's `ResidualAttentionBlock`
(TinyCLIP/src/open_clip/model.py:230,238-283): parameter names of the `nn.MultiheadAttention` it
wraps (`in_proj_weight`, `in_proj_bias`, `out_proj.weight`, `out_proj.bias`), the (length, batch,
embed) layout, an optional additive `attn_mask` (the text tower's causal mask, model.py:756-762) and
the pruning multipliers `head_z` / `hidden_z`.  The in-projection and out-projection are the sliced
GEMM (full slice), q·kᵀ + mask, softmax and ·v one kernel (`cream_attn_fwd` with its dense additive
logit term); head_dim must be 64 (ViT-B/32 image tower 12 × 64, text tower 8 × 64) and
length <= 208 (50 / 77 in the reference configs).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .autoformer.functional import DenseAttentionFn, SlicedLinearFn


class ClipAttention(nn.Module):
    def __init__(self, d_model: int, n_head: int):
        super().__init__()
        assert d_model == n_head * ops.HEAD_DIM, "fused attention kernel is built for head_dim 64"
        self.embed_dim, self.num_heads = d_model, n_head
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)

    def forward(self, x, attn_mask=None, *, head_z=None, hidden_z=None):
        """x: (length, batch, embed) as in the reference; returns the same layout."""
        L, B, E = x.shape
        H = self.num_heads
        xb = x.transpose(0, 1).contiguous()                                        # (B, L, E)
        qkv = SlicedLinearFn.apply(xb, self.in_proj_weight, self.in_proj_bias, E, 3 * E, False)
        dense = None
        if attn_mask is not None:
            if attn_mask.dtype == torch.bool:      # nn.MultiheadAttention: True = "may not attend"
                attn_mask = torch.zeros(attn_mask.shape, dtype=torch.float32, device=attn_mask.device) \
                    .masked_fill_(attn_mask, float("-inf"))
            dense = attn_mask.to(device=x.device, dtype=torch.float32).reshape(1, 1, L, L).contiguous()
        out = DenseAttentionFn.apply(qkv, H, float(ops.HEAD_DIM ** -0.5), dense)       # (B, L, E)
        if head_z is not None:
            out = (out.view(B, L, H, ops.HEAD_DIM) * head_z.view(1, 1, -1, 1).to(out.dtype)).view(B, L, E)
        out = SlicedLinearFn.apply(out, self.out_proj.weight, self.out_proj.bias, E, E, False)
        if hidden_z is not None:
            out = out * hidden_z.to(out.dtype)
        return out.transpose(0, 1)
