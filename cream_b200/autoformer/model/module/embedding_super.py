"""Drop-in for AutoFormer/model/module/embedding_super.py.

The 16x16 / stride-16 patch convolution is a GEMM over im2col patches with the first
`sample_embed_dim` filters (embedding_super.py:27-40); the module keeps the reference's `proj`
Conv2d parameter (checkpoint names / shapes) and the attributes the model reads (`num_patches`).
"""
from __future__ import annotations

import torch.nn as nn

from ..utils import to_2tuple
from ...functional import PatchEmbedFn


class PatchembedSuper(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, scale=False):
        super().__init__()
        self.img_size, self.patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        grid = tuple(i // p for i, p in zip(self.img_size, self.patch_size))
        self.num_patches = grid[0] * grid[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.super_embed_dim, self.scale = embed_dim, scale
        self.sample_embed_dim = self.sampled_weight = self.sampled_bias = self.sampled_scale = None

    def set_sample_config(self, sample_embed_dim):
        self.sample_embed_dim = sample_embed_dim
        # views for the counters below; the kernel reads the full tensors with E as an extent
        self.sampled_weight = self.proj.weight[:sample_embed_dim]
        self.sampled_bias = self.proj.bias[:sample_embed_dim]
        self.sampled_scale = self.super_embed_dim / sample_embed_dim if self.scale else None

    def forward(self, x):
        B, C, H, W = x.shape
        want = self.img_size
        assert (H, W) == tuple(want), f"Input image size ({H}*{W}) doesn't match model ({want[0]}*{want[1]})."
        assert self.patch_size[0] == self.patch_size[1] and H == W, "square patches / images only"
        y = PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, self.sample_embed_dim, self.patch_size[0])
        return y * self.sampled_scale if self.scale else y

    def calc_sampled_param_num(self):
        return int(self.sampled_weight.numel()) + int(self.sampled_bias.numel())

    def get_complexity(self, sequence_length):
        per_token = int(self.sampled_weight.numel())          # filters x (channels * patch area)
        bias_adds = self.sampled_bias.size(0) if self.sampled_bias is not None else 0
        return sequence_length * per_token + bias_adds
