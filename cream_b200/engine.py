"""Fused sampled-subnet engine: forward + backward of the AutoFormer supernet on the
B200 kernels, functional over the reference's parameter names / full-supernet shapes.

Mirrors Vision_TransformerSuper.forward after set_sample_config(config)
(AutoFormer/model/supernet_transformer.py:102-127, 147-172) and
TransformerEncoderLayer.forward (:251-287) with AttentionSuper.forward
(model/module/multihead_super.py:133-160), but as ~11 kernel launches per block:

    LN -> QKV GEMM (slice in the TMA map) -> fused attention+RPE -> proj GEMM (+bias,
    DropPath scale, residual, fp32) -> LN -> fc1 GEMM (+bias, GELU) -> fc2 GEMM (+bias,
    DropPath scale, residual)

Everything here is orchestration: device buffers come from torch, every arithmetic step
is a cream_b200 C-ABI call.  Gradients are produced as full-size fp32 tensors that are
zero outside the sampled slice, and parameters of un-sampled (identity) layers get no
gradient at all — the contract DDP(find_unused_parameters=True) relies on
(AutoFormer/supernet_train.py:288).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib, ops
from ._lib import EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_F32, EPI_F32_RESID, check

_p, _stream = ops._p, ops._stream


@dataclass
class SupernetGeometry:
    """Static (super) configuration of a Vision_TransformerSuper instance."""
    embed_dim: int
    depth: int
    num_heads: int
    mlp_ratio: float
    img_size: int = 224
    patch_size: int = 16
    in_chans: int = 3
    num_classes: int = 1000
    max_relative_position: int = 14
    gp: bool = True
    relative_position: bool = True
    abs_pos: bool = True
    eps: float = 1e-5
    qkv_bias: bool = True

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def num_tokens(self) -> int:
        return self.grid * self.grid + 1


def block_param_names(i: int, relative_position: bool = True, qkv_bias: bool = True) -> List[str]:
    p = f"blocks.{i}."
    names = [p + "attn_layer_norm.weight", p + "attn_layer_norm.bias", p + "attn.qkv.weight"]
    if qkv_bias:
        names.append(p + "attn.qkv.bias")
    if relative_position:
        names += [p + f"attn.rel_pos_embed_{kv}.embeddings_table_{vh}" for kv in "kv" for vh in "vh"]
    names += [p + "attn.proj.weight", p + "attn.proj.bias", p + "ffn_layer_norm.weight", p + "ffn_layer_norm.bias",
              p + "fc1.weight", p + "fc1.bias", p + "fc2.weight", p + "fc2.bias"]
    return names


def sampled_param_names(geo: SupernetGeometry, config: dict) -> List[str]:
    """Parameters that take part in a sampled forward (identity layers excluded)."""
    names = ["patch_embed_super.proj.weight", "patch_embed_super.proj.bias", "cls_token"]
    if geo.abs_pos:
        names.append("pos_embed")
    for i in range(config["layer_num"]):
        names += block_param_names(i, geo.relative_position, geo.qkv_bias)
    names += ["norm.weight", "norm.bias", "head.weight", "head.bias"]
    return names


def validate_config(geo: SupernetGeometry, config: dict) -> None:
    L = config["layer_num"]
    assert 1 <= L <= geo.depth, "layer_num out of range"
    assert len(config["embed_dim"]) >= L and len(config["num_heads"]) >= L and len(config["mlp_ratio"]) >= L
    E = config["embed_dim"][0]
    assert all(e == E for e in config["embed_dim"][:L]), \
        "one embed_dim per subnet (supernet_engine.py:21 replicates a single choice)"
    assert E <= geo.embed_dim and E % 4 == 0, "embed_dim must be <= super embed dim and a multiple of 4"
    for i in range(L):
        assert 64 * config["num_heads"][i] <= geo.embed_dim, "heads exceed the supernet"
        assert int(E * config["mlp_ratio"][i]) <= int(geo.embed_dim * geo.mlp_ratio)


def _tables(P, prefix, kv):
    return P[prefix + f"attn.rel_pos_embed_{kv}.embeddings_table_v"], P[prefix + f"attn.rel_pos_embed_{kv}.embeddings_table_h"]


def _pack_af(tv: torch.Tensor, th: torch.Tensor) -> torch.Tensor:
    """Two (2*max_rel+2, 64) fp32 tables -> one (1, 64, 64) bf16 pack: v rows [0,32), h rows [32,64)."""
    nb = tv.shape[0]
    pack = ops.new_pack(1, tv.device)
    ops.pack_tables(pack, 1, tv, nb, 0, (0, tv.stride(0), tv.stride(1)), th, nb, 32, (0, th.stride(0), th.stride(1)))
    return pack


class Saved:
    __slots__ = ("blocks", "cols", "x_last", "mean_f", "rstd_f", "pooled", "B", "config", "scales")


def forward(P: Dict[str, torch.Tensor], geo: SupernetGeometry, config: dict, images: torch.Tensor,
            drop_path_scales: Optional[List[Optional[torch.Tensor]]] = None, save: bool = True):
    """Sampled-subnet forward.  Returns (logits fp32 (B, num_classes), Saved|None).

    drop_path_scales: per block, None or a (2, B) fp32 tensor of per-sample DropPath factors
    (mask / keep_prob, model/utils.py:71-99) for the attention and FFN branches.
    """
    validate_config(geo, config)
    lib = _lib.load()
    dev = images.device
    assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
    B = images.shape[0]
    N, T = geo.num_tokens, geo.num_tokens - 1
    M = B * N
    E = config["embed_dim"][0]
    Es = geo.embed_dim
    sh = ops.SHADOWS
    saved = Saved() if save else None

    # ---- patch embedding as a sliced GEMM over im2col patches ----
    kdim = geo.in_chans * geo.patch_size ** 2
    cols = ops.empty_bf16(B * T, kdim, dev)
    check(lib.cream_patch_im2col(_p(images), _p(cols), cols.stride(0), B, geo.in_chans, geo.img_size, geo.img_size,
                                 geo.patch_size, _stream()), "cream_patch_im2col")
    w_pe = sh.get(P["patch_embed_super.proj.weight"])
    patch = ops.linear_fwd(cols, w_pe, E, kdim, P["patch_embed_super.proj.bias"])
    x = ops.empty_f32(M, E, dev)
    pos = P["pos_embed"] if geo.abs_pos else None
    check(lib.cream_tokens_assemble_fwd(_p(patch), patch.stride(0), _p(P["cls_token"]), _p(pos), Es, _p(x),
                                        x.stride(0), B, N, E, _stream()), "cream_tokens_assemble_fwd")

    idx = (None, None, None, None)
    if geo.relative_position:
        iv, ih, _, _ = ops.autoformer_index_tables(N, geo.max_relative_position, dev)
        idx = (iv, ih, iv, ih)
    scale = 64 ** -0.5  # (64*h // h) ** -0.5 with change_qkv, multihead_super.py:110
    af = (geo.grid, geo.max_relative_position) if geo.relative_position else None

    packs = None
    if geo.relative_position:   # every (layer, k|v) table pair packed in ONE launch
        packs = ops.pack_tables_batch([_tables(P, f"blocks.{i}.", kv) for i in range(config["layer_num"]) for kv in "kv"], dev)

    blocks = []
    for i in range(config["layer_num"]):
        pre = f"blocks.{i}."
        h = config["num_heads"][i]
        qd = 64 * h
        ffn = int(E * config["mlp_ratio"][i])
        dps = drop_path_scales[i] if drop_path_scales is not None else None
        ln1, mu1, rs1 = ops.layernorm_fwd(x, P[pre + "attn_layer_norm.weight"], P[pre + "attn_layer_norm.bias"],
                                          geo.eps, E, save_stats=save)
        wq = sh.get(P[pre + "attn.qkv.weight"], qkv=True)
        qkv = ops.qkv_fwd(ln1, wq, h, E, Es, P.get(pre + "attn.qkv.bias"))
        tk = tv = None
        if geo.relative_position:
            tk, tv = packs[2 * i:2 * i + 1], packs[2 * i + 1:2 * i + 2]
        att, lse = ops.attention_fwd(qkv, B, h, N, scale, tk=tk, tv=tv, idx=idx, need_lse=save, af=af)
        wp = sh.get(P[pre + "attn.proj.weight"])
        x1 = ops.linear_fwd(att, wp, E, qd, P[pre + "attn.proj.bias"], epi=EPI_F32_RESID, resid=x,
                            row_scale=dps[0] if dps is not None else None, rows_per_scale=N)
        ln2, mu2, rs2 = ops.layernorm_fwd(x1, P[pre + "ffn_layer_norm.weight"], P[pre + "ffn_layer_norm.bias"],
                                          geo.eps, E, save_stats=save)
        w1 = sh.get(P[pre + "fc1.weight"])
        hpre = ops.empty_bf16(M, ffn, dev)
        act = ops.linear_fwd(ln2, w1, ffn, E, P[pre + "fc1.bias"], epi=EPI_BF16_GELU, aux=hpre)
        w2 = sh.get(P[pre + "fc2.weight"])
        x2 = ops.linear_fwd(act, w2, E, ffn, P[pre + "fc2.bias"], epi=EPI_F32_RESID, resid=x1,
                            row_scale=dps[1] if dps is not None else None, rows_per_scale=N)
        if save:
            blocks.append(dict(x=x, ln1=ln1, mu1=mu1, rs1=rs1, qkv=qkv, att=att, lse=lse, tk=tk, tv=tv, x1=x1,
                               ln2=ln2, mu2=mu2, rs2=rs2, hpre=hpre, act=act, h=h, ffn=ffn))
        x = x2

    y, mu_f, rs_f = ops.layernorm_fwd(x, P["norm.weight"], P["norm.bias"], geo.eps, E, out_f32=True, save_stats=save)
    pooled = ops.empty_bf16(B, E, dev)
    first, count = (1, N - 1) if geo.gp else (0, 1)
    check(lib.cream_pool_fwd(_p(y), y.stride(0), _p(pooled), pooled.stride(0), B, N, E, first, count, _stream()),
          "cream_pool_fwd")
    w_head = sh.get(P["head.weight"])
    logits = ops.linear_fwd(pooled, w_head, geo.num_classes, E, P["head.bias"], epi=EPI_F32)
    if save:
        saved.blocks, saved.cols, saved.x_last = blocks, cols, x
        saved.mean_f, saved.rstd_f, saved.pooled = mu_f, rs_f, pooled
        saved.B, saved.config, saved.scales = B, config, drop_path_scales
    return logits[:, :geo.num_classes], saved


def backward(P: Dict[str, torch.Tensor], geo: SupernetGeometry, saved: Saved, dlogits: torch.Tensor,
             grads: Optional[Dict[str, torch.Tensor]] = None, on_group_done=None) -> Dict[str, torch.Tensor]:
    """Backward of `forward`.  Returns {name: full-size fp32 grad} for the sampled parameters
    (accumulating into `grads` when given, e.g. views of flat all-reduce buckets).
    on_group_done(name) is called with "head", "block<i>", "embed" as soon as every gradient
    of that parameter group has been enqueued (the hook point for overlapped all-reduce)."""
    lib = _lib.load()
    config = saved.config
    dev = dlogits.device
    B = saved.B
    N, T = geo.num_tokens, geo.num_tokens - 1
    M = B * N
    E = config["embed_dim"][0]
    Es = geo.embed_dim
    sh = ops.SHADOWS
    names = sampled_param_names(geo, config)
    G = grads if grads is not None else {}
    for n in names:
        if n not in G or G[n] is None:
            G[n] = torch.zeros_like(P[n], dtype=torch.float32)

    # ---- head ----
    dl = ops.empty_bf16(B, geo.num_classes, dev)
    dl.copy_(dlogits)
    ops.bias_grad(dl, G["head.bias"])
    ops.linear_wgrad(dl, saved.pooled, geo.num_classes, E, G["head.weight"])
    dpooled = ops.linear_dgrad(dl, sh.get(P["head.weight"]), geo.num_classes, E)
    dy = ops.empty_f32(M, E, dev)
    first, count = (1, N - 1) if geo.gp else (0, 1)
    check(lib.cream_pool_bwd(_p(dpooled), dpooled.stride(0), _p(dy), dy.stride(0), B, N, E, first, count, _stream()),
          "cream_pool_bwd")
    g = ops.layernorm_bwd(dy, saved.x_last, P["norm.weight"], saved.mean_f, saved.rstd_f, E, G["norm.weight"],
                          G["norm.bias"])
    if on_group_done is not None:
        on_group_done("head")

    idx = (None, None, None, None)
    if geo.relative_position:
        iv, ih, _, _ = ops.autoformer_index_tables(N, geo.max_relative_position, dev)
        idx = (iv, ih, iv, ih)
    scale = 64 ** -0.5
    af = (geo.grid, geo.max_relative_position) if geo.relative_position else None

    dpacks = None
    if geo.relative_position:
        dpacks = torch.zeros((2 * config["layer_num"], ops.NB_PACK, ops.HEAD_DIM), dtype=torch.float32, device=dev)

    for i in reversed(range(config["layer_num"])):
        pre = f"blocks.{i}."
        s = saved.blocks[i]
        h, ffn = s["h"], s["ffn"]
        qd = 64 * h
        dps = saved.scales[i] if saved.scales is not None else None
        # ---- FFN branch: x2 = x1 + s * fc2(gelu(fc1(ln2))) ----
        dy2 = ops.cast_scale(g, dps[1] if dps is not None else None, N, dbias=G[pre + "fc2.bias"])
        ops.linear_wgrad(dy2, s["act"], E, ffn, G[pre + "fc2.weight"])
        dh = ops.linear_dgrad(dy2, sh.get(P[pre + "fc2.weight"]), E, ffn, epi=EPI_BF16_DGELU, aux=s["hpre"])
        ops.bias_grad(dh, G[pre + "fc1.bias"])
        ops.linear_wgrad(dh, s["ln2"], ffn, E, G[pre + "fc1.weight"])
        dln2 = ops.linear_dgrad(dh, sh.get(P[pre + "fc1.weight"]), ffn, E)
        g1 = ops.layernorm_bwd(dln2, s["x1"], P[pre + "ffn_layer_norm.weight"], s["mu2"], s["rs2"], E,
                               G[pre + "ffn_layer_norm.weight"], G[pre + "ffn_layer_norm.bias"], resid_grad=g)
        # ---- attention branch: x1 = x + s * proj(attn(qkv(ln1))) ----
        dy1 = ops.cast_scale(g1, dps[0] if dps is not None else None, N, dbias=G[pre + "attn.proj.bias"])
        ops.linear_wgrad(dy1, s["att"], E, qd, G[pre + "attn.proj.weight"])
        datt = ops.linear_dgrad(dy1, sh.get(P[pre + "attn.proj.weight"]), E, qd)
        dqkv, _, _, _ = ops.attention_bwd(s["qkv"], s["att"], s["lse"], datt, B, h, N, scale, tk=s["tk"],
                                          tv=s["tv"], idx=idx, af=af,
                                          dtk=dpacks[2 * i:2 * i + 1] if dpacks is not None else None,
                                          dtv=dpacks[2 * i + 1:2 * i + 2] if dpacks is not None else None)
        if geo.relative_position and on_group_done is not None:
            # table gradients of this layer must be final before its bucket is reduced
            ops.unpack_table_grads_batch(dpacks[2 * i:2 * i + 2], [
                (G[pre + f"attn.rel_pos_embed_{kv}.embeddings_table_v"], G[pre + f"attn.rel_pos_embed_{kv}.embeddings_table_h"])
                for kv in "kv"])
        if geo.qkv_bias:
            ops.bias_grad(dqkv, G[pre + "attn.qkv.bias"])
        ops.qkv_wgrad(dqkv, s["ln1"], h, E, G[pre + "attn.qkv.weight"])
        dln1 = ops.qkv_dgrad(dqkv, sh.get(P[pre + "attn.qkv.weight"], qkv=True), h, E, Es)
        g = ops.layernorm_bwd(dln1, s["x"], P[pre + "attn_layer_norm.weight"], s["mu1"], s["rs1"], E,
                              G[pre + "attn_layer_norm.weight"], G[pre + "attn_layer_norm.bias"], resid_grad=g1)
        saved.blocks[i] = None  # release activations as we go
        if on_group_done is not None:
            on_group_done("block%d" % i)

    if geo.relative_position and on_group_done is None:   # single launch for all layers
        ops.unpack_table_grads_batch(dpacks, [
            (G[f"blocks.{i}.attn.rel_pos_embed_{kv}.embeddings_table_v"], G[f"blocks.{i}.attn.rel_pos_embed_{kv}.embeddings_table_h"])
            for i in range(config["layer_num"]) for kv in "kv"])

    # ---- embedding ----
    dpatch = ops.empty_bf16(B * T, E, dev)
    gpos = G["pos_embed"] if geo.abs_pos else None
    check(lib.cream_tokens_assemble_bwd(_p(g), g.stride(0), _p(dpatch), dpatch.stride(0), _p(gpos), Es,
                                        _p(G["cls_token"]), B, N, E, _stream()), "cream_tokens_assemble_bwd")
    ops.bias_grad(dpatch, G["patch_embed_super.proj.bias"])
    kdim = geo.in_chans * geo.patch_size ** 2
    ops.linear_wgrad(dpatch, saved.cols, E, kdim, G["patch_embed_super.proj.weight"])
    if on_group_done is not None:
        on_group_done("embed")
    return G


class _SupernetFn(torch.autograd.Function):
    """autograd bridge over the PYTHON sequencing of the kernels (one ctypes call per launch).  Kept as
    the readable statement of the launch sequence and as the cross-check of the native runtime, which
    issues the very same launches from C++ (tests assert bit-identical results)."""

    @staticmethod
    def forward(ctx, geo, config, names, drop_path_scales, images, *params):
        P = dict(zip(names, params))
        logits, saved = forward(P, geo, config, images, drop_path_scales, save=True)
        ctx.geo, ctx.names, ctx.saved = geo, names, saved
        ctx.save_for_backward(*params)
        return logits.contiguous()

    @staticmethod
    def backward(ctx, dlogits):
        P = dict(zip(ctx.names, ctx.saved_tensors))
        G = backward(P, ctx.geo, ctx.saved, dlogits.contiguous().float())
        ctx.saved = None
        return (None, None, None, None, None) + tuple(G[n] for n in ctx.names)


# --------------------------------------------------------------------------------------------------
# native runtime (csrc/vit_engine.cu): the default execution path
# --------------------------------------------------------------------------------------------------
USE_NATIVE = True


def native_geometry(geo: SupernetGeometry):
    from .native import VitGeometry
    return VitGeometry(embed_dim=geo.embed_dim, depth=geo.depth, num_classes=geo.num_classes, img_size=geo.img_size,
                       patch_size=geo.patch_size, in_chans=geo.in_chans, eps=geo.eps, gp=geo.gp, scale=64 ** -0.5,
                       rpe="autoformer" if geo.relative_position else "none",
                       max_relative_position=geo.max_relative_position)


def native_runner(P: Dict[str, torch.Tensor], geo: SupernetGeometry, owner=None):
    """The NativeVit bound to this parameter set (cached on `owner` when given, rebuilt if the
    parameters moved)."""
    from .native import AUTOFORMER, NativeVit
    key = tuple(P[n].data_ptr() for n in ("cls_token", "head.weight")) + (geo.num_classes,)
    cache = owner.__dict__.setdefault("_cream_native", {}) if owner is not None else {}
    r = cache.get("runner")
    if r is None or cache.get("key") != key:
        r = NativeVit(P, native_geometry(geo), AUTOFORMER)
        cache["runner"], cache["key"] = r, key
    return r


class _NativeFn(torch.autograd.Function):
    """autograd bridge over the native runtime: ONE C call for the forward, one for the backward."""

    @staticmethod
    def forward(ctx, runner, config, names, drop_path_scales, images, *params):
        runner.refresh_shadows(only_stale=True)
        logits = runner.forward(config, images, drop_path_scales)
        ctx.runner, ctx.names, ctx.generation = runner, names, runner.generation
        ctx.shapes = [(p.shape, p.device) for p in params]
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        r = ctx.runner
        if r.generation != ctx.generation:
            raise RuntimeError("cream_b200: the activations of this forward were overwritten by a later forward of the "
                               "same model (the native runtime keeps ONE set of saved activations per model); run "
                               "backward before the next forward, or set cream_b200.engine.USE_NATIVE = False")
        G = {n: torch.zeros(shape, dtype=torch.float32, device=dev) for n, (shape, dev) in zip(ctx.names, ctx.shapes)}
        r.bind_grads(G)
        dl = ops.empty_f32(dlogits.shape[0], dlogits.shape[1], dlogits.device)
        dl.copy_(dlogits)
        r.backward(dl)
        return (None, None, None, None, None) + tuple(G[n] for n in ctx.names)


def supernet_apply(P: Dict[str, torch.Tensor], geo: SupernetGeometry, config: dict, images: torch.Tensor,
                   drop_path_scales=None, owner=None) -> torch.Tensor:
    """Differentiable fused forward over a {name: parameter} dict (reference names)."""
    validate_config(geo, config)
    grad = torch.is_grad_enabled() and any(p.requires_grad for p in P.values())
    if USE_NATIVE:
        runner = native_runner(P, geo, owner)
        if not grad:
            runner.refresh_shadows(only_stale=True)
            return runner.forward(config, images, drop_path_scales).clone()
        names = sampled_param_names(geo, config)
        return _NativeFn.apply(runner, config, names, drop_path_scales, images, *[P[n] for n in names])
    if not grad:
        return forward(P, geo, config, images, drop_path_scales, save=False)[0].contiguous()
    names = sampled_param_names(geo, config)
    return _SupernetFn.apply(geo, config, names, drop_path_scales, images, *[P[n] for n in names])
