"""Drop-in for AutoFormer/model/module/qkv_super.py.

The reference rebuilds the sampled weight on every `set_sample_config` by concatenating three
strided row selections of the full tensor (rows i, i+3, i+6, ... for i = 0, 1, 2;
qkv_super.py:45-51,72-77) and takes the bias as a plain prefix (:80-83).  Here the bf16 shadow of
the weight is stored de-interleaved once per optimizer step, so a sampled (in, 3*64*heads) slice is
three dense row blocks of one TMA tensor map; `samples['weight']` keeps the reference's SHAPE for the
parameter / FLOP counters (an `(out, in)` prefix view has the same element count as the concatenation).
"""
from __future__ import annotations

from .Linear_super import _SlicedLinear


class qkv_super(_SlicedLinear):
    _interleaved_qkv = True
    _init_on_construct = False    # qkv_super.py:21 leaves the nn.Linear default initialisation in place
