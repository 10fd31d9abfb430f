"""Supernet training step on the fused engine, with data-parallel gradient all-reduce.

`SupernetTrainer.step` is the body of the reference's hot loop
(AutoFormer/supernet_engine.py:49-107): host batch -> device, sample a subnet
(`sample_configs`, :13-24, identical on every rank because the Python RNG is seeded with the
epoch, :36), set_sample_config, forward, loss, backward, optimizer step.  Differences that
are B200-first rather than a translation:

  * the whole sampled subnet is two engine calls (forward / backward), no autograd graph over
    the model and no DDP graph walk for unused parameters: gradients land directly in views of
    per-layer flat fp32 buckets and un-sampled layers simply keep `grad = None`
    (the find_unused_parameters=True contract of supernet_train.py:288);
  * each layer's bucket is all-reduced (NCCL over NVLink/NVSwitch, average) on a side stream as
    soon as that layer's backward has been enqueued, overlapping the remaining backward;
  * bf16 compute with fp32 masters instead of fp16 autocast + GradScaler (no loss scaling,
    no per-step inf check / host sync);
  * no host synchronisation inside a step unless the caller reads the loss.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import engine, ops
from .autoformer.model.supernet_transformer import Vision_TransformerSuper


def sample_configs(choices: dict, rnd=random) -> dict:
    """AutoFormer/supernet_engine.py:13-24 (same draw order from the same RNG)."""
    config = {}
    depth = rnd.choice(choices['depth'])
    for dimension in ['mlp_ratio', 'num_heads']:
        config[dimension] = [rnd.choice(choices[dimension]) for _ in range(depth)]
    config['embed_dim'] = [rnd.choice(choices['embed_dim'])] * depth
    config['layer_num'] = depth
    return config


class GradBuckets:
    """Per-layer flat fp32 gradient buckets; parameters' .grad are views into them."""

    def __init__(self, model: Vision_TransformerSuper):
        named = dict(model.named_parameters())
        self.groups: Dict[str, List[str]] = {"embed": [], "head": []}
        for name in named:
            if name.startswith("blocks."):
                self.groups.setdefault("block%d" % int(name.split(".")[1]), []).append(name)
            elif name in ("norm.weight", "norm.bias", "head.weight", "head.bias"):
                self.groups["head"].append(name)
            else:
                self.groups["embed"].append(name)
        self.flat: Dict[str, torch.Tensor] = {}
        self.views: Dict[str, torch.Tensor] = {}
        # one master buffer (a single memset per step), one contiguous 256-byte aligned span per group
        # (the unit of all-reduce), parameters at 16-byte aligned offsets inside it
        sizes = {g: (sum((named[n].numel() + 3) // 4 * 4 for n in names) + 63) // 64 * 64
                 for g, names in self.groups.items()}
        dev = next(iter(named.values())).device
        self.master = torch.zeros(sum(sizes.values()), dtype=torch.float32, device=dev)
        base = 0
        self.end_of: Dict[str, int] = {}       # group -> end / start offset of its span in the master buffer
        self.start_of: Dict[str, int] = {}
        for gname, names in self.groups.items():
            buf = self.master[base:base + sizes[gname]]
            base += sizes[gname]
            self.end_of[gname] = base
            self.start_of[gname] = base - sizes[gname]
            off = 0
            for n in names:
                p = named[n]
                self.views[n] = buf[off:off + p.numel()].view(p.shape)
                off += (p.numel() + 3) // 4 * 4
            self.flat[gname] = buf

    def group_of_layer(self, i: int) -> str:
        return "block%d" % i


class StagedBatch:
    """A batch whose host->device copy is in flight on the trainer's copy stream."""
    __slots__ = ("images", "targets", "ready", "slot")

    def __init__(self, images, targets, ready, slot=-1):
        self.images, self.targets, self.ready, self.slot = images, targets, ready, slot


class SupernetTrainer:
    def __init__(self, model: Vision_TransformerSuper, choices: dict, lr: float = 5e-4, weight_decay: float = 0.05,
                 process_group=None, native: bool = True, overlap=1):
        """native=True (default): the step is a handful of C calls - cream_vit_fwd, cream_xent_fwd_bwd,
        cream_vit_bwd (one call per all-reduce group when world > 1), cream_adamw_step (AdamW fused with
        the bf16 shadow refresh).  native=False: the same kernels sequenced from Python with torch's
        fused AdamW and torch's cross-entropy (kept as the cross-check of the native runtime)."""
        if not model.fusable():
            raise ValueError("SupernetTrainer drives the fused engine: the model must be pre-norm, change_qkv=True, "
                             "scale=False, drop_rate=0, attn_drop_rate=0 (DropPath is supported); use the module "
                             "path under a stock torch loop otherwise")
        self.model = model
        self.choices = choices
        self.overlap = overlap
        self.geo = model.engine_geometry()
        self.params = dict(model.named_parameters())
        self.buckets = GradBuckets(model)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1:
            # what DDP's constructor does (supernet_train.py:288): replicas start from rank 0's weights and
            # buffers, whatever each rank's seed was (the reference seeds with args.seed + rank, :197-198)
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
        # timm create_optimizer('adamw') (supernet_train.py:296) -> add_weight_decay: no decay on 1-D
        # params, biases and names that are EXACTLY in no_weight_decay() — so 'rel_pos_embed' matches
        # nothing and the 2-D relative-position tables ARE decayed, as in the reference recipe; torch's
        # fused AdamW is the same update rule on device.
        skip = model.no_weight_decay()
        decay, no_decay = [], []
        for n, p in self.params.items():
            (no_decay if (p.ndim <= 1 or n.endswith(".bias") or n in skip) else decay).append(n)
        self.native = None
        if native and model.pos_embed.is_cuda:
            from .native import FlatAdamW
            self.native = engine.native_runner(self.params, self.geo, owner=model)   # shared with model(x)
            self.native.bind_grads(self.buckets.views)
            qkv_names = {f"blocks.{i}.attn.qkv.weight" for i in range(self.geo.depth)}
            self.optimizer = FlatAdamW(self.params, self.buckets.views, set(decay), lr, weight_decay,
                                       shadows=self.native.shadows, qkv_interleaved_names=qkv_names)
            d = max(choices["depth"])
            self.native.reserve({"layer_num": d, "embed_dim": [max(choices["embed_dim"])] * d,
                                 "num_heads": [max(choices["num_heads"])] * d,
                                 "mlp_ratio": [max(choices["mlp_ratio"])] * d}, 1)
        else:
            self.optimizer = torch.optim.AdamW([{"params": [self.params[n] for n in decay], "weight_decay": weight_decay},
                                                {"params": [self.params[n] for n in no_decay], "weight_decay": 0.0}],
                                               lr=lr, fused=model.pos_embed.is_cuda)
        self._sampled_cache: Dict[int, tuple] = {}
        self.comm_stream = torch.cuda.Stream() if (self.world > 1 and model.pos_embed.is_cuda) else None
        self.copy_stream = torch.cuda.Stream() if model.pos_embed.is_cuda else None
        # two persistent device staging slots (no allocator traffic, hence no cudaMalloc, in steady state)
        self._slots = [None, None]
        self._slot_free = [None, None]      # event: the step that consumed the slot has read it
        self._next_slot = 0
        self.last_config: Optional[dict] = None

    # ------------------------------------------------------------------------------------
    def _allreduce(self, gname: str):
        if self.world == 1:
            return
        buf = self.buckets.flat[gname]
        if self.comm_stream is None:          # CPU / gloo (host-logic tests)
            dist.all_reduce(buf, group=self.pg)
            buf.div_(self.world)
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.pg)

    def _sampled(self, cfg) -> tuple:
        """Names of the parameters that take part in a subnet of this depth (identity layers excluded)."""
        L = cfg["layer_num"]
        if L not in self._sampled_cache:
            self._sampled_cache[L] = tuple(engine.sampled_param_names(self.geo, cfg))
        return self._sampled_cache[L]

    def _assign_grads(self, names) -> None:
        sampled = set(names)
        for n, p in self.params.items():
            p.grad = self.buckets.views[n] if n in sampled else None

    def _backward(self, saved, dlogits):
        """engine.backward (Python sequencing) with per-layer bucket all-reduce interleaved."""
        cfg = saved.config
        names = self._sampled(cfg)
        self.buckets.master.zero_()          # one memset (140 MB for supernet-S, ~25 us) instead of one per group
        G = {n: self.buckets.views[n] for n in names}
        engine.backward(self.params, self.geo, saved, dlogits, grads=G,
                        on_group_done=self._allreduce if self.world > 1 else None)
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._assign_grads(names)

    def _backward_native(self, cfg, dlogits):
        """cream_vit_bwd in ONE C call, then ONE all-reduce over the contiguous span of the sampled
        groups (embed | head | block 0 .. L-1 sit at the front of the master buffer; un-sampled layers
        behind them are neither zero-filled by kernels nor reduced).

        `overlap` = number of collectives per step.  1 (default): the single all-reduce above.  k > 1:
        k - 1 collectives over the upper blocks run on a side stream under the remaining backward and
        the last one - the prefix of the buffer - follows it.  True: one collective per group as in round 1.
        Measured at N = 2 (profiles/r02_scaling.md): NCCL's resident CTAs take SMs away from the
        persistent, statically scheduled GEMM / attention kernels (+8 % on every kernel that runs
        beside a collective): weak-scaling efficiency 0.978 with ONE exposed exchange (0.27 ms of a
        12.3 ms step), 0.972 with two collectives, 0.969 with three, 0.93 with one per layer."""
        self.buckets.master.zero_()
        if self.native.G is not self.buckets.views:      # an autograd call through model(x) re-bound them
            self.native.bind_grads(self.buckets.views)
        L = cfg["layer_num"]
        if self.world == 1:
            self.native.backward(dlogits)
            return
        chunks = max(1, int(self.overlap) if self.overlap is not True else L + 2)
        if chunks == 1:
            self.native.backward(dlogits)
            span = self.buckets.master[:self.buckets.end_of["block%d" % (L - 1)]]
            dist.all_reduce(span, op=dist.ReduceOp.AVG, group=self.pg)
            return
        if chunks >= L + 2:          # one collective per group (round 1's schedule)
            self.native.backward(dlogits, 0, 0)
            self._allreduce("head")
            for stage in range(1, L + 1):
                self.native.backward(dlogits, stage, stage)
                self._allreduce("block%d" % (L - stage))
            self.native.backward(dlogits, L + 1, L + 1)
            self._allreduce("embed")
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            return
        # `chunks` collectives: chunks - 1 of them cover the upper blocks in equal shares and run on the side
        # stream under the rest of the backward; the last one (embed | head | lowest blocks: a prefix of the
        # master buffer) follows the backward on the compute stream.
        per = max(1, L // chunks)
        bounds = [L - per * c for c in range(chunks)]           # block index where each overlapped chunk starts
        hi, stage_done = L, 0
        for lo in bounds[1:]:
            self.native.backward(dlogits, stage_done, L - lo)   # stages up to and including layer `lo`
            stage_done = L - lo + 1
            span = self.buckets.master[self.buckets.start_of["block%d" % lo]:self.buckets.end_of["block%d" % (hi - 1)]]
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(span, op=dist.ReduceOp.AVG, group=self.pg)
            hi = lo
        self.native.backward(dlogits, stage_done, L + 1)
        dist.all_reduce(self.buckets.master[:self.buckets.end_of["block%d" % (hi - 1)]], op=dist.ReduceOp.AVG, group=self.pg)
        torch.cuda.current_stream().wait_stream(self.comm_stream)

    def stage(self, images: torch.Tensor, targets: torch.Tensor) -> StagedBatch:
        """Start the host->device copy of the NEXT batch on a side stream and return a handle for
        `step`.  With pinned host tensors (the reference's DataLoader(pin_memory=True) +
        `.to(device, non_blocking=True)`, supernet_engine.py:57-58) the copy is a DMA that runs under
        the kernels of the step in flight instead of in front of its own step."""
        dev = self.model.pos_embed.device
        if self.copy_stream is None or images.is_cuda:
            return StagedBatch(images.to(dev), targets.to(dev), None)
        i = self._next_slot
        self._next_slot ^= 1
        slot = self._slots[i]
        if slot is None or slot[0].shape != images.shape or slot[0].dtype != images.dtype or \
                slot[1].shape != targets.shape or slot[1].dtype != targets.dtype:
            slot = (torch.empty(images.shape, dtype=images.dtype, device=dev),
                    torch.empty(targets.shape, dtype=targets.dtype, device=dev))
            self._slots[i] = slot
            self._slot_free[i] = None
            self.copy_stream.wait_stream(torch.cuda.current_stream())
            for t in slot:                      # allocated on the compute stream, written on the copy stream
                t.record_stream(self.copy_stream)
        with torch.cuda.stream(self.copy_stream):
            if self._slot_free[i] is not None:
                self.copy_stream.wait_event(self._slot_free[i])     # the previous user of the slot has read it
            slot[0].copy_(images, non_blocking=True)
            slot[1].copy_(targets, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        return StagedBatch(slot[0], slot[1], ready, i)

    def forward_backward(self, images, targets: Optional[torch.Tensor] = None, config: Optional[dict] = None,
                         rnd=random) -> torch.Tensor:
        """Sample (or take) a subnet, run forward, cross-entropy and backward with the overlapped
        gradient all-reduce; parameters' `.grad` hold the (rank-averaged) gradients afterwards.
        Returns the (device) loss tensor of the LOCAL batch without synchronising."""
        model = self.model
        dev = model.pos_embed.device
        staged = images if isinstance(images, StagedBatch) else None
        if staged is not None:
            if staged.ready is not None:
                torch.cuda.current_stream().wait_event(staged.ready)
            images, targets = staged.images, staged.targets
        else:
            images = images.to(dev, non_blocking=True)
            targets = targets.to(dev, non_blocking=True)
        cfg = config if config is not None else sample_configs(self.choices, rnd)
        self.last_config = cfg
        model.set_sample_config(cfg)
        scales = model.drop_path_scales(images.shape[0], dev)
        images = images.float().contiguous()
        if self.native is not None:
            engine.validate_config(self.geo, cfg)
            logits = self.native.forward(cfg, images, scales)
            loss, dlogits = self.native.xent(logits, targets)
            loss = loss[0]
        else:
            logits, saved = engine.forward(self.params, self.geo, cfg, images, scales, save=True)
            lg = logits.detach().requires_grad_(True)
            loss = F.cross_entropy(lg, targets)
            (dlogits,) = torch.autograd.grad(loss, lg)
        if staged is not None and staged.slot >= 0:      # images (im2col) and targets (loss) have been consumed
            ev = torch.cuda.Event()
            ev.record()
            self._slot_free[staged.slot] = ev
        if self.native is not None:
            self._backward_native(cfg, dlogits)
        else:
            self._backward(saved, dlogits)
        return loss.detach()

    def step(self, images, targets: Optional[torch.Tensor] = None, config: Optional[dict] = None,
             rnd=random) -> torch.Tensor:
        """One training step (the loop body of supernet_engine.py:49-107); returns the (device) loss
        tensor without synchronising.  `images` is a tensor (host or device) or a StagedBatch."""
        loss = self.forward_backward(images, targets, config, rnd)
        if self.native is not None:
            self.optimizer.step(self._sampled(self.last_config), cache_key=self.last_config["layer_num"])
            self.native.mark_shadows_fresh()
        else:
            self.optimizer.step()
            # torch's fused AdamW updates the parameters WITHOUT bumping their version counters, so the
            # version-tagged shadow cache cannot see the update (round 1 trained on stale bf16 weights
            # because of exactly this): drop the tags explicitly
            ops.SHADOWS.invalidate()
        return loss

    def sync_grads_to_params(self) -> None:
        """Expose the gradients of the last step as `p.grad` (views of the flat buckets; None for the
        parameters of un-sampled layers).  The native step does not need it; tests and external
        optimizers do."""
        self._assign_grads(self._sampled(self.last_config))
