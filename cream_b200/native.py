"""Python face of the native runtime (csrc/vit_engine.cu, csrc/optim.cu).

`NativeVit` binds a {name: parameter} dict (the reference's names and FULL supernet shapes) to a
`cream_vit_desc` once — parameter, shadow and gradient pointers never change — and then runs a whole
sampled-subnet forward or backward with ONE C call each: no per-launch Python, no allocation (every
activation lives in a persistent arena), TMA descriptors cached by address.  Two parameter layouts:

  * `AUTOFORMER`  Vision_TransformerSuper (AutoFormer/model/supernet_transformer.py): interleaved
                  qkv rows, 2-D relative position tables on K and V, gp pooling;
  * `DEIT_IRPE`   the DeiT VisionTransformer with iRPE (iRPE/DeiT-with-iRPE/rpe_vision_transformer.py,
                  contextual product tables on keys, shared head): plain [q; k; v] rows, cls pooling.

`FlatAdamW` is the optimizer of the fused path: torch.optim.AdamW semantics over every parameter in one
launch, writing the bf16 weight shadows in the same pass.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, ops
from ._lib import AdamwSeg, VitDesc, check

_p, _stream = ops._p, ops._stream


@dataclass
class Layout:
    """Parameter names of one model family (format strings take the block index)."""
    patch: str
    cls: str = "cls_token"
    pos: str = "pos_embed"
    ln1: str = ""
    qkv: str = ""
    proj: str = ""
    ln2: str = ""
    fc1: str = ""
    fc2: str = ""
    norm: str = "norm"
    head: str = "head"
    tables: tuple = ()               # up to 4 names: K-side pair, V-side pair ("" = absent)
    qkv_interleaved: bool = True


AUTOFORMER = Layout(
    patch="patch_embed_super.proj", ln1="blocks.{i}.attn_layer_norm", qkv="blocks.{i}.attn.qkv",
    proj="blocks.{i}.attn.proj", ln2="blocks.{i}.ffn_layer_norm", fc1="blocks.{i}.fc1", fc2="blocks.{i}.fc2",
    tables=("blocks.{i}.attn.rel_pos_embed_k.embeddings_table_v", "blocks.{i}.attn.rel_pos_embed_k.embeddings_table_h",
            "blocks.{i}.attn.rel_pos_embed_v.embeddings_table_v", "blocks.{i}.attn.rel_pos_embed_v.embeddings_table_h"),
    qkv_interleaved=True)

DEIT_IRPE = Layout(
    patch="patch_embed.proj", ln1="blocks.{i}.norm1", qkv="blocks.{i}.attn.qkv", proj="blocks.{i}.attn.proj",
    ln2="blocks.{i}.norm2", fc1="blocks.{i}.mlp.fc1", fc2="blocks.{i}.mlp.fc2",
    tables=("blocks.{i}.attn.rpe_k.lookup_table_weight", "", "", ""), qkv_interleaved=False)


@dataclass
class VitGeometry:
    """Static description of the (super) network the parameters belong to."""
    embed_dim: int                     # E*: width of the full tensors
    depth: int
    num_classes: int = 1000
    img_size: int = 224
    patch_size: int = 16
    in_chans: int = 3
    eps: float = 1e-5
    gp: bool = True
    scale: float = 0.125
    rpe: str = "autoformer"            # "autoformer" | "irpe_product" | "none"
    max_relative_position: int = 14
    irpe: tuple = (1.9, 1)             # (ratio, skip) for rpe == "irpe_product"

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def num_tokens(self) -> int:
        return self.grid * self.grid + 1


def _shadow_buffer(w: torch.Tensor) -> torch.Tensor:
    rows = w.shape[0]
    cols = w.numel() // rows
    return torch.empty((rows, ops.round_up(cols, 8)), dtype=torch.bfloat16, device=w.device)


class NativeVit:
    def __init__(self, params: Dict[str, torch.Tensor], geo: VitGeometry, layout: Layout = AUTOFORMER,
                 grads: Optional[Dict[str, torch.Tensor]] = None):
        self.lib = _lib.load()
        self.P, self.geo, self.layout = params, geo, layout
        some = next(iter(params.values()))
        self.device = some.device
        assert some.is_cuda, "cream_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path"
        self.G: Dict[str, torch.Tensor] = grads if grads is not None else {}
        self.shadows: Dict[str, torch.Tensor] = {}
        self.desc = VitDesc()
        self.arena: Optional[torch.Tensor] = None
        self.logits: Optional[torch.Tensor] = None
        self.generation = 0
        self._keep: List[torch.Tensor] = []
        self._bind_static()

    # ------------------------------------------------------------------------------------------
    def _w(self, name):           # weight names that own a bf16 shadow
        return name + ".weight"

    def weight_names(self) -> List[str]:
        lay, names = self.layout, []
        names.append(self._w(lay.patch))
        for i in range(self.geo.depth):
            names += [self._w(t.format(i=i)) for t in (lay.qkv, lay.proj, lay.fc1, lay.fc2)]
        names.append(self._w(lay.head))
        return names

    def _bind_static(self):
        d, P, geo, lay = self.desc, self.P, self.geo, self.layout
        for n in self.weight_names():
            self.shadows[n] = _shadow_buffer(P[n])
        d.N, d.num_classes, d.in_chans = geo.num_tokens, geo.num_classes, geo.in_chans
        d.img_size, d.patch_size, d.eps = geo.img_size, geo.patch_size, geo.eps
        d.pool_first, d.pool_count = (1, geo.num_tokens - 1) if geo.gp else (0, 1)
        d.qkv_interleaved = int(lay.qkv_interleaved)
        wq = P[self._w(lay.qkv.format(i=0))]
        d.qkv_group_rows = wq.shape[0] // 3
        sh = lambda t: self.shadows[self._w(t.format(i=0) if "{i}" in t else t)]
        full = lambda t: P[self._w(t.format(i=0) if "{i}" in t else t)]
        cols = lambda w: w.numel() // w.shape[0]
        d.ld_wqkv, d.ld_wproj = sh(lay.qkv).stride(0), sh(lay.proj).stride(0)
        d.ld_wfc1, d.ld_wfc2 = sh(lay.fc1).stride(0), sh(lay.fc2).stride(0)
        d.ld_wpatch, d.ld_whead = sh(lay.patch).stride(0), sh(lay.head).stride(0)
        d.ld_gqkv, d.ld_gproj, d.ld_gfc1 = cols(full(lay.qkv)), cols(full(lay.proj)), cols(full(lay.fc1))
        d.ld_gfc2, d.ld_gpatch, d.ld_ghead = cols(full(lay.fc2)), cols(full(lay.patch)), cols(full(lay.head))
        d.scale = geo.scale
        # relative-position gather tables
        N = geo.num_tokens
        if geo.rpe == "autoformer":
            iv, ih, _, _ = ops.autoformer_index_tables(N, geo.max_relative_position, self.device)
            d.idx_a, d.idx_b, d.idx_va, d.idx_vb, d.ld_idx = _p(iv), _p(ih), _p(iv), _p(ih), iv.stride(0)
            d.af_grid, d.af_max_rel = geo.grid, geo.max_relative_position
            t0 = P[lay.tables[0].format(i=0)]
            d.tab_nb, d.tab_row_off1 = t0.shape[0], 32
            d.tab_stride_b, d.tab_stride_d = t0.stride(0), t0.stride(1)
            d.tabv_stride_b, d.tabv_stride_d = t0.stride(0), t0.stride(1)
        elif geo.rpe == "irpe_product":
            ratio, skip = geo.irpe
            ids, nb = ops.irpe_bucket_ids(3, geo.grid, geo.grid, skip, 1 * ratio, 2 * ratio, 8 * ratio)
            it = ops.irpe_index_table_u8(ids, self.device)
            d.idx_a, d.ld_idx = _p(it), it.stride(0)
            gp = ops.irpe_grid_product_structure(ids, geo.grid, skip)
            if gp is not None:
                d.gp_grid, d.gp_w, d.gp_skip_id = geo.grid, gp[0], gp[1]
                C.memmove(d.gp_lut_a, gp[2].ctypes.data, 32)
                C.memmove(d.gp_lut_b, gp[3].ctypes.data, 32)
            t0 = P[lay.tables[0].format(i=0)]              # (1, 64, nb): [0, channel, bucket]
            assert t0.shape[0] == 1 and t0.shape[1] == ops.HEAD_DIM and t0.shape[2] == nb, "shared-head contextual table"
            d.tab_nb, d.tab_row_off1 = nb, 0
            d.tab_stride_b, d.tab_stride_d = t0.stride(2), t0.stride(1)
            d.tabv_stride_b, d.tabv_stride_d = t0.stride(2), t0.stride(1)
        # embedding / head
        d.wpatch, d.bpatch = _p(self.shadows[self._w(lay.patch)]), _p(P.get(lay.patch + ".bias"))
        d.cls = _p(P[lay.cls])
        pos = P.get(lay.pos)
        d.pos, d.ld_pos = _p(pos), (pos.shape[-1] if pos is not None else 0)
        d.norm_g, d.norm_b = _p(P[lay.norm + ".weight"]), _p(P[lay.norm + ".bias"])
        d.whead, d.bhead = _p(self.shadows[self._w(lay.head)]), _p(P.get(lay.head + ".bias"))
        for i in range(geo.depth):
            L = d.layers[i]
            f = lambda t: t.format(i=i)
            L.ln1_g, L.ln1_b = _p(P[f(lay.ln1) + ".weight"]), _p(P[f(lay.ln1) + ".bias"])
            L.ln2_g, L.ln2_b = _p(P[f(lay.ln2) + ".weight"]), _p(P[f(lay.ln2) + ".bias"])
            for key, t in (("qkv", lay.qkv), ("proj", lay.proj), ("fc1", lay.fc1), ("fc2", lay.fc2)):
                setattr(L, "w" + key, _p(self.shadows[self._w(f(t))]))
                setattr(L, "b" + key, _p(P.get(f(t) + ".bias")))
            for k, t in enumerate(lay.tables):
                L.tab[k] = _p(P[f(t)]) if (t and geo.rpe != "none") else None
        self.refresh_shadows()
        if self.G:
            self.bind_grads(self.G)

    def mark_shadows_fresh(self) -> None:
        """The shadows were just rewritten by FlatAdamW: adopt the parameters' current versions."""
        seen = self.__dict__.setdefault("_versions", {})
        for n in self.shadows:
            seen[n] = (self.P[n]._version, self.P[n].data_ptr())

    def bind_grads(self, G: Dict[str, torch.Tensor]) -> None:
        """Point the descriptor at full-size fp32 gradient tensors (accumulated into, caller-zeroed)."""
        self.G = G
        d, lay, geo = self.desc, self.layout, self.geo
        g = lambda n: _p(G.get(n))
        d.g_wpatch, d.g_bpatch = g(lay.patch + ".weight"), g(lay.patch + ".bias")
        d.g_cls, d.g_pos = g(lay.cls), g(lay.pos)
        d.g_norm_g, d.g_norm_b = g(lay.norm + ".weight"), g(lay.norm + ".bias")
        d.g_whead, d.g_bhead = g(lay.head + ".weight"), g(lay.head + ".bias")
        for i in range(geo.depth):
            L = d.layers[i]
            f = lambda t: t.format(i=i)
            L.g_ln1_g, L.g_ln1_b = g(f(lay.ln1) + ".weight"), g(f(lay.ln1) + ".bias")
            L.g_ln2_g, L.g_ln2_b = g(f(lay.ln2) + ".weight"), g(f(lay.ln2) + ".bias")
            for key, t in (("qkv", lay.qkv), ("proj", lay.proj), ("fc1", lay.fc1), ("fc2", lay.fc2)):
                setattr(L, "g_w" + key, g(f(t) + ".weight"))
                setattr(L, "g_b" + key, g(f(t) + ".bias"))
            for k, t in enumerate(lay.tables):
                L.g_tab[k] = g(f(t)) if (t and geo.rpe != "none") else None

    def refresh_shadows(self, only_stale: bool = False) -> None:
        """fp32 -> bf16 cast of the GEMM weights (construction, load_state_dict, updates by a torch
        optimizer); in steady state FlatAdamW writes the shadows itself.  only_stale: recast just the
        weights whose tensor version changed since their last cast."""
        lay = self.layout
        qkv_names = {self._w(lay.qkv.format(i=i)) for i in range(self.geo.depth)}
        seen = self.__dict__.setdefault("_versions", {})
        for n, sh in self.shadows.items():
            w = self.P[n].detach()
            tag = (self.P[n]._version, self.P[n].data_ptr())
            if only_stale and seen.get(n) == tag:
                continue
            seen[n] = tag
            w2 = w.reshape(w.shape[0], -1)
            if n in qkv_names and lay.qkv_interleaved:
                check(self.lib.cream_shadow_qkv(_p(w2), _p(sh), w2.shape[0] // 3, w2.shape[1], w2.stride(0), sh.stride(0),
                                                _stream()), "cream_shadow_qkv")
            else:
                check(self.lib.cream_shadow_cast(_p(w2), _p(sh), w2.shape[0], w2.shape[1], w2.stride(0), sh.stride(0),
                                                 _stream()), "cream_shadow_cast")

    # ------------------------------------------------------------------------------------------
    def _configure(self, config: dict, batch: int):
        d = self.desc
        E = config["embed_dim"][0]
        L = config["layer_num"]
        assert all(e == E for e in config["embed_dim"][:L]), "one embed_dim per subnet (supernet_engine.py:21)"
        assert 1 <= L <= self.geo.depth and E <= self.geo.embed_dim and E % 4 == 0
        d.B, d.E, d.depth = batch, E, L
        for i in range(L):
            d.layers[i].heads = config["num_heads"][i]
            d.layers[i].ffn = int(E * config["mlp_ratio"][i])
        need = self.lib.cream_vit_arena_bytes(C.byref(d))
        if need < 0:
            raise _lib.CreamError("cream_vit_arena_bytes: invalid configuration (see stderr)")
        if self.arena is None or self.arena.numel() < need:
            self.arena = None
            self.arena = torch.empty(int(need * 1.02) + 4096, dtype=torch.uint8, device=self.device)
        base = self.arena.data_ptr()
        d.arena = (base + 255) // 256 * 256
        d.arena_bytes = self.arena.numel() - (d.arena - base)
        if self.logits is None or self.logits.shape[0] < batch:
            self.logits = ops.empty_f32(batch, self.geo.num_classes, self.device)
        d.logits, d.ld_logits = _p(self.logits), self.logits.stride(0)

    def reserve(self, config: dict, batch: int) -> None:
        """Size the arena for (the largest) `config` up front so that no step ever allocates."""
        self._configure(config, batch)

    def forward(self, config: dict, images: torch.Tensor, drop_path_scales=None) -> torch.Tensor:
        """Sampled-subnet forward; returns the persistent fp32 logits buffer (B, num_classes)."""
        assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
        B = images.shape[0]
        self._configure(config, B)
        d = self.desc
        d.images = _p(images)
        for i in range(d.depth):
            s = drop_path_scales[i] if drop_path_scales is not None else None
            d.layers[i].dp_scale = _p(s)
        self._keep = [images] + ([s for s in drop_path_scales if s is not None] if drop_path_scales else [])
        self.generation += 1
        rc = self.lib.cream_vit_fwd(C.byref(d), _stream())
        check(rc, "cream_vit_fwd", kernels=self.lib.cream_vit_last_launches())
        return self.logits[:B, :self.geo.num_classes]

    def backward(self, dlogits: torch.Tensor, first_stage: int = 0, last_stage: Optional[int] = None) -> None:
        """Backward stages [first_stage, last_stage] (0 = head, 1..depth = layers depth-1..0,
        depth+1 = embedding) of the LAST forward; gradients accumulate into the bound tensors."""
        d = self.desc
        assert dlogits.dtype == torch.float32 and dlogits.stride(1) == 1 and dlogits.stride(0) % 4 == 0
        d.dlogits, d.ld_dlogits = _p(dlogits), dlogits.stride(0)
        last = d.depth + 1 if last_stage is None else last_stage
        rc = self.lib.cream_vit_bwd(C.byref(d), first_stage, last, _stream())
        check(rc, "cream_vit_bwd", kernels=self.lib.cream_vit_last_launches())

    def xent(self, logits: torch.Tensor, targets: torch.Tensor):
        """Mean cross-entropy and its gradient w.r.t. the logits in one kernel: (loss (1,), dlogits)."""
        B, Cn = logits.shape
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        dl = ops.empty_f32(B, Cn, logits.device)
        assert targets.dtype == torch.int64 and targets.is_cuda
        check(self.lib.cream_xent_fwd_bwd(_p(logits), logits.stride(0), _p(targets), _p(loss), _p(dl), dl.stride(0), B, Cn,
                                          _stream()), "cream_xent_fwd_bwd")
        return loss, dl


class FlatAdamW:
    """torch.optim.AdamW semantics (decoupled decay, per-parameter step counts, parameters without a
    gradient skipped) over all parameters in ONE launch, fused with the bf16 shadow refresh."""

    def __init__(self, named_params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], decay_names: set,
                 lr: float, weight_decay: float, betas=(0.9, 0.999), eps: float = 1e-8,
                 shadows: Optional[Dict[str, torch.Tensor]] = None, qkv_interleaved_names: Optional[set] = None):
        self.lib = _lib.load()
        self.names = list(named_params)
        self.index = {n: i for i, n in enumerate(self.names)}
        self.lr, self.betas, self.eps = lr, betas, eps
        dev = next(iter(named_params.values())).device
        total = sum((p.numel() + 3) // 4 * 4 for p in named_params.values())      # every state slice 16-byte aligned
        self.state = torch.zeros(2 * total, dtype=torch.float32, device=dev)      # exp_avg | exp_avg_sq
        segs = (AdamwSeg * len(self.names))()
        off = 0
        self.max_numel = 1
        for i, n in enumerate(self.names):
            p, g = named_params[n], grads[n]
            assert p.is_contiguous() and g.is_contiguous() and p.dtype == torch.float32
            s = segs[i]
            s.p, s.g = p.data_ptr(), g.data_ptr()
            s.m = self.state.data_ptr() + 4 * off
            s.v = self.state.data_ptr() + 4 * (total + off)
            s.numel = p.numel()
            s.weight_decay = weight_decay if n in decay_names else 0.0
            s.step = 0
            sh = shadows.get(n) if shadows else None
            if sh is not None:
                s.shadow, s.rows, s.cols, s.shadow_ld = sh.data_ptr(), p.shape[0], p.numel() // p.shape[0], sh.stride(0)
                s.qkv_group_rows = p.shape[0] // 3 if (qkv_interleaved_names and n in qkv_interleaved_names) else 0
            else:
                s.shadow, s.rows, s.cols, s.shadow_ld, s.qkv_group_rows = None, 1, max(p.numel(), 1), 0, 0
            off += (p.numel() + 3) // 4 * 4
            self.max_numel = max(self.max_numel, p.numel())
        raw = np.frombuffer(bytes(segs), dtype=np.uint8).copy()
        self.segs_dev = torch.from_numpy(raw).to(dev)
        chunk = self.lib.cream_adamw_chunk()
        blocks = [(i, first) for i, n in enumerate(self.names) for first in range(0, named_params[n].numel(), chunk)]
        self.n_blocks = len(blocks)
        self.blocks_dev = torch.tensor(blocks, dtype=torch.int32).to(dev)
        self._masks: Dict[object, torch.Tensor] = {}
        self._plists: Dict[object, list] = {}
        self._params = named_params

    def step(self, active_names, cache_key=None) -> None:
        """One AdamW step over the parameters named in `active_names` (those that received a gradient
        this step).  `cache_key` (e.g. the sampled depth) lets repeated active sets reuse their device mask."""
        key = cache_key if cache_key is not None else tuple(active_names)
        mask = self._masks.get(key)
        if mask is None:
            a = np.zeros(len(self.names), dtype=np.int32)
            a[[self.index[n] for n in active_names]] = 1
            mask = torch.from_numpy(a).to(self.segs_dev.device)
            self._masks[key] = mask
        b1, b2 = self.betas
        check(self.lib.cream_adamw_step(_p(self.segs_dev), _p(mask), len(self.names), _p(self.blocks_dev), self.n_blocks,
                                        self.lr, b1, b2, self.eps, _stream()), "cream_adamw_step", kernels=2)
        # the kernel wrote the parameters through raw pointers: tell autograd / the shadow caches
        plist = self._plists.get(key)
        if plist is None:
            plist = self._plists[key] = [self._params[n] for n in active_names]
        torch.autograd.graph.increment_version(plist)

    def steps_taken(self) -> Dict[str, int]:
        """Per-parameter step counts (device -> host; for tests and checkpoints)."""
        raw = self.segs_dev.cpu().numpy().tobytes()
        segs = (AdamwSeg * len(self.names)).from_buffer_copy(raw)
        return {n: int(segs[i].step) for i, n in enumerate(self.names)}
