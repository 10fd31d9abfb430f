"""Drop-in for iRPE/DeiT-with-iRPE/rpe_ops/rpe_index.py: the autograd wrapper of the native gather.

    Y[b, h, i, j] = input[b, h, i, index[i, j]]          (forward)
    grad_input[b, h, i, index[i, j]] += grad_Y[b, h, i, j]   (backward, into a zeroed buffer)

`irpe.py:8-15` imports `RPEIndexFunction` from here and selects int32 bucket ids when it is present
(`irpe.py:563-565`).  The extension module underneath is cream_b200's `rpe_index_cpp` (same five entry
points as the reference's pybind module, over the C ABI)."""
from __future__ import annotations

import torch

from . import rpe_index_cpp as _ext

EXPECTED_VERSION = "1.2.0"
if _ext.version() != EXPECTED_VERSION:
    raise AssertionError(f"Unmatched `rpe_index_cpp` version: {_ext.version()}, expected version: {EXPECTED_VERSION}")

_ENTRY = {("cpu", "fwd"): _ext.forward_cpu, ("cuda", "fwd"): _ext.forward_gpu,
          ("cpu", "bwd"): _ext.backward_cpu, ("cuda", "bwd"): _ext.backward_gpu}


def _entry(tensor: torch.Tensor, which: str):
    return _ENTRY[("cpu" if tensor.device.type == "cpu" else "cuda", which)]


class RPEIndexFunction(torch.autograd.Function):
    """Same contract as the reference class: `apply(input (B, H, Lq, nb), index (Lq, Lk) int32)`."""

    @staticmethod
    def forward(ctx, input, index):
        ctx.input_shape = tuple(input.shape)
        ctx.save_for_backward(index)
        return _entry(input, "fwd")(input, index)

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.needs_input_grad[0]:
            return None, None
        (index,) = ctx.saved_tensors
        grad_input = grad_output.new_zeros(ctx.input_shape)
        _entry(grad_output, "bwd")(grad_input, grad_output, index)
        return grad_input, None
