"""Drop-in for AutoFormer/model/module/layernorm_super.py: LayerNorm over the first
`sample_embed_dim` features with the matching prefix of weight and bias, as one kernel."""
from __future__ import annotations

import torch

from ...functional import SlicedLayerNormFn
from ._sliced import SliceViews


class LayerNormSuper(SliceViews, torch.nn.LayerNorm):
    def __init__(self, super_embed_dim):
        torch.nn.LayerNorm.__init__(self, super_embed_dim)
        self._init_views()
        self.super_embed_dim = super_embed_dim
        self.sample_embed_dim = None

    def set_sample_config(self, sample_embed_dim):
        self.sample_embed_dim = sample_embed_dim
        self._sample_parameters()

    def _build_views(self):
        e = self.sample_embed_dim
        return {'weight': self.weight[:e], 'bias': self.bias[:e]}

    def forward(self, x):
        self.sample_parameters()
        return SlicedLayerNormFn.apply(x, self.weight, self.bias, self.sample_embed_dim, self.eps)

    def get_complexity(self, sequence_length):
        return sequence_length * self.sample_embed_dim
