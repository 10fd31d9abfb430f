"""Helpers with the names AutoFormer/model/supernet_transformer.py imports from model.utils
(trunc_normal_, DropPath, to_2tuple) — AutoFormer/model/utils.py:49-107."""
from __future__ import annotations

import collections.abc

import torch
import torch.nn as nn


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """Truncated normal initialiser (same contract as model/utils.py:49-67)."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _tuple_of(n: int):
    """x -> (x,) * n unless x is already iterable (the reference's to_2tuple family, model/utils.py:8-17)."""
    return lambda x: x if isinstance(x, collections.abc.Iterable) else (x,) * n


to_1tuple, to_2tuple, to_3tuple, to_4tuple = (_tuple_of(k) for k in (1, 2, 3, 4))
to_ntuple = _tuple_of


def drop_path_scale(batch: int, drop_prob: float, training: bool, device) -> torch.Tensor | None:
    """Per-sample factor floor(keep + U[0,1)) / keep of model/utils.py:71-87, as a (B,) fp32
    tensor that the GEMM epilogue applies (None when DropPath is inactive)."""
    if drop_prob == 0. or not training:
        return None
    keep = 1.0 - drop_prob
    return torch.floor(keep + torch.rand(batch, dtype=torch.float32, device=device)) / keep


def drop_path(x, drop_prob: float = 0., training: bool = False):
    if drop_prob == 0. or not training:
        return x
    s = drop_path_scale(x.shape[0], drop_prob, training, x.device).to(x.dtype)
    return x * s.view((x.shape[0],) + (1,) * (x.ndim - 1))


class DropPath(nn.Module):
    """Stochastic depth per sample (model/utils.py:90-99)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)
