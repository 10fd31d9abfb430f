"""Module with the surface of the reference extension `rpe_index_cpp`
(rpe_ops/rpe_index.cpp:126-142): version / forward_gpu / backward_gpu over the cream_b200
C ABI.  There is deliberately no CPU implementation: forward_cpu / backward_cpu raise."""
from __future__ import annotations

import sys

from cream_b200 import _lib, ops


def version() -> str:
    return _lib.load().cream_rpe_index_version().decode()


def forward_gpu(input, index):
    return ops.rpe_index_forward(input, index)


def backward_gpu(grad_input, grad_output, index):
    ops.rpe_index_backward(grad_input, grad_output, index)


def forward_cpu(input, index):
    raise RuntimeError("cream_b200 rpe_index: CPU tensors are not supported (sm_100a only, no fallback)")


def backward_cpu(grad_input, grad_output, index):
    raise RuntimeError("cream_b200 rpe_index: CPU tensors are not supported (sm_100a only, no fallback)")


# `import rpe_index_cpp` (top level) is what rpe_ops/rpe_index.py:2 of the reference does.
sys.modules.setdefault("rpe_index_cpp", sys.modules[__name__])
