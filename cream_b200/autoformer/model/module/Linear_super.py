"""Drop-in for AutoFormer/model/module/Linear_super.py.

`LinearSuper` keeps the reference's constructor, attributes (`super_*_dim`, `sample_*_dim`,
`samples`, `sample_scale`) and methods, and still IS an `nn.Linear` (the model's `_init_weights`
dispatches on that).  The forward is the sliced tcgen05 GEMM: the top-left (out, in) rectangle
of the full weight is described to the kernel by extents, never copied; there is no CPU path.
`_SlicedLinear` is the part shared with qkv_super.
"""
from __future__ import annotations

import torch.nn as nn

from ...functional import SlicedLinearFn
from ._sliced import SliceViews, xavier_or


class _SlicedLinear(SliceViews, nn.Linear):
    _interleaved_qkv = False      # qkv_super: rows i, i+3, i+6, ... form the q / k / v blocks
    _init_on_construct = True     # the reference initialises LinearSuper but not qkv_super

    def __init__(self, super_in_dim, super_out_dim, bias=True, uniform_=None, non_linear='linear', scale=False):
        nn.Linear.__init__(self, super_in_dim, super_out_dim, bias=bias)
        self._init_views()
        self.super_in_dim, self.super_out_dim = super_in_dim, super_out_dim
        self.sample_in_dim = self.sample_out_dim = None
        self.scale = scale
        if self._init_on_construct:
            self._reset_parameters(bias, uniform_, non_linear)

    def _reset_parameters(self, bias, uniform_, non_linear):
        xavier_or(self.weight, self.bias if bias else None, uniform_, non_linear)

    def set_sample_config(self, sample_in_dim, sample_out_dim):
        self.sample_in_dim, self.sample_out_dim = sample_in_dim, sample_out_dim
        self._sample_parameters()

    def _build_views(self):
        rows, cols = self.sample_out_dim, self.sample_in_dim
        self.sample_scale = self.super_out_dim / rows
        return {'weight': self.weight[:rows, :cols],
                'bias': None if self.bias is None else self.bias[:rows]}

    def forward(self, x):
        self.sample_parameters()
        y = SlicedLinearFn.apply(x, self.weight, self.bias, self.sample_in_dim, self.sample_out_dim,
                                 self._interleaved_qkv)
        return y * (self.sample_scale if self.scale else 1)

    def get_complexity(self, sequence_length):
        return sequence_length * self.sample_out_dim * self.sample_in_dim


class LinearSuper(_SlicedLinear):
    pass
