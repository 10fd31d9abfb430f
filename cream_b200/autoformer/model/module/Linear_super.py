"""Drop-in for AutoFormer/model/module/Linear_super.py: same class, constructor, attributes
and methods; forward runs the sliced tcgen05 GEMM (no sliced weight copy, no CPU path)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .... import ops
from ...functional import SlicedLinearFn


class LinearSuper(nn.Linear):
    def __init__(self, super_in_dim, super_out_dim, bias=True, uniform_=None, non_linear='linear', scale=False):
        super().__init__(super_in_dim, super_out_dim, bias=bias)
        self.super_in_dim = super_in_dim      # largest network
        self.super_out_dim = super_out_dim
        self.sample_in_dim = None             # current sampled sizes
        self.sample_out_dim = None
        self.samples = {}
        self.scale = scale
        self._reset_parameters(bias, uniform_, non_linear)   # Linear_super.py:21,32-36
        self.profiling = False

    def profile(self, mode=True):
        self.profiling = mode

    def sample_parameters(self, resample=False):
        if self.profiling or resample:
            return self._sample_parameters()
        return self.samples

    def _reset_parameters(self, bias, uniform_, non_linear):
        nn.init.xavier_uniform_(self.weight) if uniform_ is None else uniform_(self.weight, non_linear=non_linear)
        if bias:
            nn.init.constant_(self.bias, 0.)

    def set_sample_config(self, sample_in_dim, sample_out_dim):
        self.sample_in_dim = sample_in_dim
        self.sample_out_dim = sample_out_dim
        self._sample_parameters()

    def _sample_parameters(self):
        # views only (Linear_super.py:71-81); kept for calc_sampled_param_num / get_complexity
        self.samples['weight'] = self.weight[:self.sample_out_dim, :self.sample_in_dim]
        self.samples['bias'] = self.bias
        self.sample_scale = self.super_out_dim / self.sample_out_dim
        if self.bias is not None:
            self.samples['bias'] = self.bias[:self.sample_out_dim]
        return self.samples

    def forward(self, x):
        self.sample_parameters()
        y = SlicedLinearFn.apply(x, self.weight, self.bias, self.sample_in_dim, self.sample_out_dim, False)
        return y * (self.sample_scale if self.scale else 1)

    def calc_sampled_param_num(self):
        assert 'weight' in self.samples.keys()
        weight_numel = self.samples['weight'].numel()
        bias_numel = self.samples['bias'].numel() if self.samples['bias'] is not None else 0
        return weight_numel + bias_numel

    def get_complexity(self, sequence_length):
        return sequence_length * np.prod(self.samples['weight'].size())
