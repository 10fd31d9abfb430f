"""Double-buffered host -> device input staging shared by the trainers.

The reference feeds its models from DataLoader(pin_memory=True) + `.to(device, non_blocking=True)`
(AutoFormer/supernet_engine.py:57-58, DeiT-with-iRPE/engine.py, TinyCLIP/src/training/train.py); issued on the
compute stream that copy sits in front of its own step.  `BatchStager.stage` starts the copy of the NEXT batch on
a side stream into one of two persistent device slots, so the DMA runs under the kernels of the step in flight.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class Staged:
    """A batch whose copy is in flight: `.tensors` are device tensors valid once `.acquire()` returned."""
    __slots__ = ("tensors", "ready", "slot", "_owner")

    def __init__(self, tensors, ready, slot, owner):
        self.tensors, self.ready, self.slot, self._owner = tensors, ready, slot, owner

    def acquire(self):
        """Make the current stream wait for the copy; returns the device tensors."""
        if self.ready is not None:
            torch.cuda.current_stream().wait_event(self.ready)
        return self.tensors

    def release(self):
        """Call after the LAST kernel that reads the tensors has been enqueued on the current stream."""
        if self.slot is not None:
            ev = torch.cuda.Event()
            ev.record()
            self._owner._free[self.slot] = ev


class BatchStager:
    def __init__(self, device):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots: List[Optional[List[torch.Tensor]]] = [None, None]
        self._free: List[Optional[torch.cuda.Event]] = [None, None]
        self._next = 0

    def stage(self, *host: torch.Tensor) -> Staged:
        if all(t.is_cuda for t in host):
            return Staged(list(host), None, None, self)
        i = self._next
        self._next ^= 1
        slot = self._slots[i]
        if slot is None or any(s.shape != h.shape or s.dtype != h.dtype for s, h in zip(slot, host)) or len(slot) != len(host):
            slot = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host]
            self._slots[i], self._free[i] = slot, None
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            for t in slot:              # allocated on the compute stream, written on the copy stream
                t.record_stream(self.copy_stream)
        with torch.cuda.stream(self.copy_stream):
            if self._free[i] is not None:
                self.copy_stream.wait_event(self._free[i])      # the previous user of the slot has read it
            for s, h in zip(slot, host):
                s.copy_(h, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        return Staged(slot, ready, i, self)
