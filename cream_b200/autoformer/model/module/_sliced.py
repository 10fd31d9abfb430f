"""What every weight-entangled drop-in shares.

The reference modules (AutoFormer/model/module/*.py) each carry their own copy of the same
bookkeeping: a `samples` dict of sliced parameter views, a `profiling` switch that forces the
views to be rebuilt, and parameter / FLOP counters read by the evolution search.  Here it lives
once.  Nothing in this file touches the GPU: the views exist only so that
`calc_sampled_param_num`, `get_complexity` and external readers of `module.samples[...]` see
what they see with the reference; the kernels take the FULL tensors plus the sampled extents.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class SliceViews:
    """Mixin: owns `samples` / `profiling` and the reference's accessor names."""

    samples: Dict[str, Optional[torch.Tensor]]

    def _init_views(self) -> None:
        self.samples = {}
        self.profiling = False

    # -- to be provided by the concrete module ------------------------------------------------
    def _build_views(self) -> Dict[str, Optional[torch.Tensor]]:
        raise NotImplementedError

    # -- the reference's method names ---------------------------------------------------------
    def profile(self, mode=True):
        self.profiling = mode

    def _sample_parameters(self):
        self.samples.update(self._build_views())
        return self.samples

    def sample_parameters(self, resample=False):
        rebuild = resample or self.profiling
        return self._sample_parameters() if rebuild else self.samples

    def calc_sampled_param_num(self):
        views = [v for v in self.samples.values() if v is not None]
        if not views:
            raise AssertionError("set_sample_config() has not been called")
        return sum(int(v.numel()) for v in views)


def xavier_or(weight: torch.Tensor, bias: Optional[torch.Tensor], uniform_, non_linear) -> None:
    """Initialisation of the reference's sliced linears (Linear_super.py:32-36): Xavier-uniform
    unless the caller supplies its own `uniform_(weight, non_linear=...)`, zero bias."""
    if uniform_ is None:
        torch.nn.init.xavier_uniform_(weight)
    else:
        uniform_(weight, non_linear=non_linear)
    if bias is not None:
        torch.nn.init.constant_(bias, 0.)
