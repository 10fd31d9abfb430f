"""Drop-in for iRPE/DeiT-with-iRPE/rpe_ops/rpe_index.py (RPEIndexFunction)."""
from __future__ import annotations

import torch

from . import rpe_index_cpp

EXPECTED_VERSION = "1.2.0"
assert rpe_index_cpp.version() == EXPECTED_VERSION, \
    f"Unmatched `rpe_index_cpp` version: {rpe_index_cpp.version()}, expected version: {EXPECTED_VERSION}"


class RPEIndexFunction(torch.autograd.Function):
    '''Y[b, h, i, j] = input[b, h, i, index[i, j]]   (rpe_index.py:11-56)'''

    @staticmethod
    def forward(ctx, input, index):
        ctx.save_for_backward(index)
        ctx.input_shape = input.shape
        fn = rpe_index_cpp.forward_cpu if input.device.type == 'cpu' else rpe_index_cpp.forward_gpu
        return fn(input, index)

    @staticmethod
    def backward(ctx, grad_output):
        index = ctx.saved_tensors[0]
        if ctx.needs_input_grad[0]:
            grad_input = grad_output.new_zeros(ctx.input_shape)
            fn = rpe_index_cpp.backward_cpu if grad_output.device.type == 'cpu' else rpe_index_cpp.backward_gpu
            fn(grad_input, grad_output, index)
            return grad_input, None
        return None, None
