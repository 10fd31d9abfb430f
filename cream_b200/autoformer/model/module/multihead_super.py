"""Drop-in for AutoFormer/model/module/multihead_super.py.

`AttentionSuper`: sliced QKV GEMM -> ONE fused kernel (q k^T, 2-D relative-position gather on keys
and values, softmax, P v) -> sliced proj GEMM.  `RelativePosition2D_super` keeps the reference's
parameters (`embeddings_table_v/h`, checkpoint names) and interface; the kernel consumes the tables
directly and never builds the (N, N, head_dim) embedding tensor of multihead_super.py:61-66.
"""
from __future__ import annotations

import torch
from torch import nn

from .... import ops
from ...functional import AutoformerAttentionFn
from ..utils import trunc_normal_
from .Linear_super import LinearSuper
from .qkv_super import qkv_super


class RelativePosition2D_super(nn.Module):
    """Two learnable tables of 2*max_relative_position + 2 rows: row 0 is the cls-token bucket,
    rows 1.. the clipped vertical (`_v`) / horizontal (`_h`) offsets (multihead_super.py:16-26)."""

    def __init__(self, num_units, max_relative_position):
        super().__init__()
        self.num_units, self.max_relative_position = num_units, max_relative_position
        rows = 2 * max_relative_position + 2
        for axis in ("v", "h"):
            table = nn.Parameter(torch.randn(rows, num_units))
            trunc_normal_(table, std=.02)
            setattr(self, f"embeddings_table_{axis}", table)
        self.sample_head_dim = None
        self.sample_embeddings_table_v = self.sample_embeddings_table_h = None

    def set_sample_config(self, sample_head_dim):
        self.sample_head_dim = sample_head_dim
        self.sample_embeddings_table_v = self.embeddings_table_v[:, :sample_head_dim]
        self.sample_embeddings_table_h = self.embeddings_table_h[:, :sample_head_dim]

    def calc_sampled_param_num(self):
        return sum(int(t.numel()) for t in (self.sample_embeddings_table_v, self.sample_embeddings_table_h))

    def index_tables(self, length_q, device):
        """(idx_v, idx_h) int64 (N, N) of multihead_super.py:40-59, built by the library."""
        _, _, iv, ih = ops.autoformer_index_tables(length_q, self.max_relative_position, device)
        return torch.from_numpy(iv).long(), torch.from_numpy(ih).long()

    def forward(self, length_q, length_k):
        """Materialised (N, N, head_dim) embeddings — interface parity only; the fused attention
        kernel never builds this tensor."""
        assert length_q == length_k
        where = self.embeddings_table_v.device
        iv, ih = (t.to(where) for t in self.index_tables(length_q, where))
        return self.sample_embeddings_table_v[iv] + self.sample_embeddings_table_h[ih]


class AttentionSuper(nn.Module):
    def __init__(self, super_embed_dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.,
                 normalization=False, relative_position=False, num_patches=None, max_relative_position=14,
                 scale=False, change_qkv=False):
        super().__init__()
        super_head_dim = super_embed_dim // num_heads
        self.num_heads, self.super_embed_dim = num_heads, super_embed_dim
        self.scale = qk_scale or super_head_dim ** -0.5
        self.fc_scale, self.change_qkv = scale, change_qkv
        self.relative_position, self.max_relative_position = relative_position, max_relative_position
        qkv_cls = qkv_super if change_qkv else LinearSuper
        self.qkv = qkv_cls(super_embed_dim, 3 * super_embed_dim, bias=qkv_bias)
        if relative_position:
            self.rel_pos_embed_k = RelativePosition2D_super(super_head_dim, max_relative_position)
            self.rel_pos_embed_v = RelativePosition2D_super(super_head_dim, max_relative_position)
        self.proj = LinearSuper(super_embed_dim, super_embed_dim)
        self.attn_drop, self.proj_drop = nn.Dropout(attn_drop), nn.Dropout(proj_drop)
        self.sample_qk_embed_dim = self.sample_v_embed_dim = None
        self.sample_num_heads = self.sample_scale = self.sample_in_embed_dim = None

    def set_sample_config(self, sample_q_embed_dim=None, sample_num_heads=None, sample_in_embed_dim=None):
        self.sample_in_embed_dim, self.sample_num_heads = sample_in_embed_dim, sample_num_heads
        if self.change_qkv:      # q/k/v width follows the sampled head count (64 per head in the supernets)
            width, scale_from = sample_q_embed_dim, sample_q_embed_dim
        else:                    # full-width q/k/v; the softmax scale follows the sampled embed dim
            width, scale_from = self.super_embed_dim, sample_in_embed_dim
        self.sample_qk_embed_dim = width
        self.sample_scale = (scale_from // sample_num_heads) ** -0.5
        self.qkv.set_sample_config(sample_in_dim=sample_in_embed_dim, sample_out_dim=3 * width)
        self.proj.set_sample_config(sample_in_dim=width, sample_out_dim=sample_in_embed_dim)
        if self.relative_position:
            for side in (self.rel_pos_embed_k, self.rel_pos_embed_v):
                side.set_sample_config(width // sample_num_heads)

    def calc_sampled_param_num(self):
        return 0      # the children (qkv, proj, tables) report their own

    def get_complexity(self, sequence_length):
        n, w = sequence_length, self.sample_qk_embed_dim
        flops = self.qkv.get_complexity(n) + self.proj.get_complexity(n)
        flops += 2 * n * n * w                                  # q k^T and P v
        if self.relative_position:                              # the reference's own estimate, both sides
            r = self.max_relative_position
            flops += r * n * n + n * n / 2.0
            flops += r * n * n + n * w / 2.0
        return flops

    def _tables(self):
        if not self.relative_position:
            return ()
        k, v = self.rel_pos_embed_k, self.rel_pos_embed_v
        return (k.embeddings_table_v, k.embeddings_table_h, v.embeddings_table_v, v.embeddings_table_h)

    def forward(self, x):
        if not self.change_qkv:
            raise NotImplementedError(
                "cream_b200 supports change_qkv=True only (every published AutoFormer command passes --change_qk; "
                "without it the reference reshape at multihead_super.py:135 is invalid for sampled heads)")
        if self.training and self.attn_drop.p != 0.0:
            raise NotImplementedError("attention dropout > 0 is not supported by the fused kernel")
        head_dim = self.sample_qk_embed_dim // self.sample_num_heads
        assert head_dim == ops.HEAD_DIM, "fused attention kernel is built for head_dim 64"
        qkv = self.qkv(x)                                           # (B, N, 3*64h) bf16
        out = AutoformerAttentionFn.apply(qkv, self.sample_num_heads, float(self.sample_scale),
                                          self.max_relative_position, *self._tables())   # (B, N, 64h)
        if self.fc_scale:
            out = out * (self.super_embed_dim / self.sample_qk_embed_dim)
        return self.proj_drop(self.proj(out))
