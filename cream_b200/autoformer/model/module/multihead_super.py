"""Drop-in for AutoFormer/model/module/multihead_super.py: AttentionSuper runs QKV GEMM ->
fused attention+RPE kernel -> proj GEMM; RelativePosition2D_super keeps the reference's
parameters and interface (its tables are consumed directly by the fused kernel)."""
from __future__ import annotations

import torch
from torch import nn

from .Linear_super import LinearSuper
from .qkv_super import qkv_super
from ..utils import trunc_normal_
from ...functional import AutoformerAttentionFn
from .... import ops


class RelativePosition2D_super(nn.Module):
    def __init__(self, num_units, max_relative_position):
        super().__init__()
        self.num_units = num_units
        self.max_relative_position = max_relative_position
        # index 0 of each table is the encoding of the cls token (multihead_super.py:21-23)
        self.embeddings_table_v = nn.Parameter(torch.randn(max_relative_position * 2 + 2, num_units))
        self.embeddings_table_h = nn.Parameter(torch.randn(max_relative_position * 2 + 2, num_units))
        trunc_normal_(self.embeddings_table_v, std=.02)
        trunc_normal_(self.embeddings_table_h, std=.02)
        self.sample_head_dim = None
        self.sample_embeddings_table_h = None
        self.sample_embeddings_table_v = None

    def set_sample_config(self, sample_head_dim):
        self.sample_head_dim = sample_head_dim
        self.sample_embeddings_table_h = self.embeddings_table_h[:, :sample_head_dim]
        self.sample_embeddings_table_v = self.embeddings_table_v[:, :sample_head_dim]

    def calc_sampled_param_num(self):
        return self.sample_embeddings_table_h.numel() + self.sample_embeddings_table_v.numel()

    def index_tables(self, length_q, device):
        """(idx_v, idx_h) int64 (N, N) of multihead_super.py:40-59, built by the library."""
        _, _, iv, ih = ops.autoformer_index_tables(length_q, self.max_relative_position, device)
        return torch.from_numpy(iv).long(), torch.from_numpy(ih).long()

    def forward(self, length_q, length_k):
        """Materialised (N, N, head_dim) embeddings — interface parity only; the fused attention
        kernel never builds this tensor."""
        assert length_q == length_k
        iv, ih = self.index_tables(length_q, self.embeddings_table_v.device)
        iv, ih = iv.to(self.embeddings_table_v.device), ih.to(self.embeddings_table_v.device)
        return self.sample_embeddings_table_v[iv] + self.sample_embeddings_table_h[ih]


class AttentionSuper(nn.Module):
    def __init__(self, super_embed_dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.,
                 normalization=False, relative_position=False, num_patches=None, max_relative_position=14,
                 scale=False, change_qkv=False):
        super().__init__()
        self.num_heads = num_heads
        head_dim = super_embed_dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.super_embed_dim = super_embed_dim
        self.fc_scale = scale
        self.change_qkv = change_qkv
        if change_qkv:
            self.qkv = qkv_super(super_embed_dim, 3 * super_embed_dim, bias=qkv_bias)
        else:
            self.qkv = LinearSuper(super_embed_dim, 3 * super_embed_dim, bias=qkv_bias)
        self.relative_position = relative_position
        if self.relative_position:
            self.rel_pos_embed_k = RelativePosition2D_super(super_embed_dim // num_heads, max_relative_position)
            self.rel_pos_embed_v = RelativePosition2D_super(super_embed_dim // num_heads, max_relative_position)
        self.max_relative_position = max_relative_position
        self.sample_qk_embed_dim = None
        self.sample_v_embed_dim = None
        self.sample_num_heads = None
        self.sample_scale = None
        self.sample_in_embed_dim = None
        self.proj = LinearSuper(super_embed_dim, super_embed_dim)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj_drop = nn.Dropout(proj_drop)

    def set_sample_config(self, sample_q_embed_dim=None, sample_num_heads=None, sample_in_embed_dim=None):
        self.sample_in_embed_dim = sample_in_embed_dim
        self.sample_num_heads = sample_num_heads
        if not self.change_qkv:
            self.sample_qk_embed_dim = self.super_embed_dim
            self.sample_scale = (sample_in_embed_dim // self.sample_num_heads) ** -0.5
        else:
            self.sample_qk_embed_dim = sample_q_embed_dim
            self.sample_scale = (self.sample_qk_embed_dim // self.sample_num_heads) ** -0.5
        self.qkv.set_sample_config(sample_in_dim=sample_in_embed_dim, sample_out_dim=3 * self.sample_qk_embed_dim)
        self.proj.set_sample_config(sample_in_dim=self.sample_qk_embed_dim, sample_out_dim=sample_in_embed_dim)
        if self.relative_position:
            self.rel_pos_embed_k.set_sample_config(self.sample_qk_embed_dim // sample_num_heads)
            self.rel_pos_embed_v.set_sample_config(self.sample_qk_embed_dim // sample_num_heads)

    def calc_sampled_param_num(self):
        return 0

    def get_complexity(self, sequence_length):
        total_flops = 0
        total_flops += self.qkv.get_complexity(sequence_length)
        total_flops += sequence_length * sequence_length * self.sample_qk_embed_dim
        total_flops += sequence_length * sequence_length * self.sample_qk_embed_dim
        total_flops += self.proj.get_complexity(sequence_length)
        if self.relative_position:
            total_flops += self.max_relative_position * sequence_length * sequence_length + sequence_length * sequence_length / 2.0
            total_flops += self.max_relative_position * sequence_length * sequence_length + sequence_length * self.sample_qk_embed_dim / 2.0
        return total_flops

    def forward(self, x):
        B, N, C = x.shape
        if not self.change_qkv:
            raise NotImplementedError(
                "cream_b200 supports change_qkv=True only (every published AutoFormer command passes --change_qk; "
                "without it the reference reshape at multihead_super.py:135 is invalid for sampled heads)")
        if self.attn_drop.p != 0.0 and self.training:
            raise NotImplementedError("attention dropout > 0 is not supported by the fused kernel")
        head_dim = self.sample_qk_embed_dim // self.sample_num_heads
        assert head_dim == ops.HEAD_DIM, "fused attention kernel is built for head_dim 64"
        qkv = self.qkv(x)                                           # (B, N, 3*64h) bf16
        tabs = ()
        if self.relative_position:
            tabs = (self.rel_pos_embed_k.embeddings_table_v, self.rel_pos_embed_k.embeddings_table_h,
                    self.rel_pos_embed_v.embeddings_table_v, self.rel_pos_embed_v.embeddings_table_h)
        out = AutoformerAttentionFn.apply(qkv, self.sample_num_heads, float(self.sample_scale),
                                          self.max_relative_position, *tabs)   # (B, N, 64h)
        if self.fc_scale:
            out = out * (self.super_embed_dim / self.sample_qk_embed_dim)
        out = self.proj(out)
        out = self.proj_drop(out)
        return out
