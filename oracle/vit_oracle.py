"""ORACLE (test infrastructure only — never imported by the product path).

Plain PyTorch fp32 CPU restatement of the floating-point part of the hot path:

  * the AutoFormer sampled-subnet forward (weight-entangled slicing, 2D relative-position
    tables on K and V, pre-norm blocks, gp pooling, head), functional over a state_dict
    with the reference's parameter names and full-supernet shapes;
  * the DeiT + iRPE attention (contextual product method on q / k / v, bias mode).

Backward comes from autograd over these functions.  Each function cites the reference
lines it follows (paths relative to the reference checkout).  Pinned against the
reference itself through tests/golden/ (see tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from . import rel_index


@dataclass
class SupernetSpec:
    """Constructor arguments of Vision_TransformerSuper as used by supernet_train.py:255-265."""
    embed_dim: int = 256
    depth: int = 14
    num_heads: int = 4
    mlp_ratio: float = 4.0
    img_size: int = 224
    patch_size: int = 16
    in_chans: int = 3
    num_classes: int = 1000
    max_relative_position: int = 14
    gp: bool = True
    relative_position: bool = True
    abs_pos: bool = True
    qkv_bias: bool = True
    head_dim: int = 64  # supernet_transformer.py:243 hard-codes 64 * heads under change_qkv

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def num_tokens(self) -> int:
        return self.grid * self.grid + 1


SUPERNET_T = SupernetSpec(256, 14, 4, 4.0)   # experiments/supernet/supernet-T.yaml
SUPERNET_S = SupernetSpec(448, 14, 7, 4.0)   # experiments/supernet/supernet-S.yaml
SUPERNET_B = SupernetSpec(640, 16, 10, 4.0)  # experiments/supernet/supernet-B.yaml

SEARCH_SPACE = {  # experiments/supernet/supernet-{T,S,B}.yaml SEARCH_SPACE
    "T": dict(mlp_ratio=[3.5, 4.0], num_heads=[3, 4], depth=[12, 13, 14], embed_dim=[192, 216, 240]),
    "S": dict(mlp_ratio=[3.0, 3.5, 4.0], num_heads=[5, 6, 7], depth=[12, 13, 14], embed_dim=[320, 384, 448]),
    "B": dict(mlp_ratio=[3.0, 3.5, 4.0], num_heads=[9, 10], depth=[14, 15, 16], embed_dim=[528, 576, 624]),
}


def sample_configs(choices: dict, rnd) -> dict:
    """AutoFormer/supernet_engine.py:13-24; `rnd` is a random.Random (the engine seeds the
    global `random` with the epoch, supernet_engine.py:36)."""
    config = {}
    depth = rnd.choice(choices["depth"])
    for dimension in ["mlp_ratio", "num_heads"]:
        config[dimension] = [rnd.choice(choices[dimension]) for _ in range(depth)]
    config["embed_dim"] = [rnd.choice(choices["embed_dim"])] * depth
    config["layer_num"] = depth
    return config


def param_shapes(spec: SupernetSpec) -> dict:
    """Names and full-supernet shapes of Vision_TransformerSuper.state_dict()
    (supernet_transformer.py:21-78, 182-222; multihead_super.py:16-26, 69-98)."""
    E = spec.embed_dim
    shapes = {
        "pos_embed": (1, spec.num_tokens, E),
        "cls_token": (1, 1, E),
        "patch_embed_super.proj.weight": (E, spec.in_chans, spec.patch_size, spec.patch_size),
        "patch_embed_super.proj.bias": (E,),
    }
    ffn = int(spec.mlp_ratio * E)
    hd = E // spec.num_heads
    nrel = spec.max_relative_position * 2 + 2
    for i in range(spec.depth):
        p = f"blocks.{i}."
        shapes[p + "attn.qkv.weight"] = (3 * E, E)
        shapes[p + "attn.qkv.bias"] = (3 * E,)
        for kv in "kv":
            for vh in "vh":
                shapes[p + f"attn.rel_pos_embed_{kv}.embeddings_table_{vh}"] = (nrel, hd)
        shapes[p + "attn.proj.weight"] = (E, E)
        shapes[p + "attn.proj.bias"] = (E,)
        for ln in ("attn_layer_norm", "ffn_layer_norm"):
            shapes[p + ln + ".weight"] = (E,)
            shapes[p + ln + ".bias"] = (E,)
        shapes[p + "fc1.weight"] = (ffn, E)
        shapes[p + "fc1.bias"] = (ffn,)
        shapes[p + "fc2.weight"] = (E, ffn)
        shapes[p + "fc2.bias"] = (E,)
    shapes["norm.weight"] = (E,)
    shapes["norm.bias"] = (E,)
    shapes["head.weight"] = (spec.num_classes, E)
    shapes["head.bias"] = (spec.num_classes,)
    return shapes


def init_params(spec: SupernetSpec, seed: int = 0, std: float = 0.02, dtype=torch.float32) -> dict:
    """Deterministic synthetic parameters (numpy PCG64, platform independent).  NOT the
    reference initialiser: biases and LayerNorm get non-trivial values on purpose so that
    slicing mistakes cannot hide behind zeros/ones."""
    import numpy as np

    g = np.random.default_rng(seed)
    out = {}
    for name, shape in param_shapes(spec).items():
        a = g.standard_normal(shape, dtype=np.float32)
        if name.endswith("layer_norm.weight") or name == "norm.weight":
            a = 1.0 + 0.1 * a
        elif name.endswith(".bias"):
            a = 0.05 * a
        else:
            a = std * a
        out[name] = torch.from_numpy(a).to(dtype)
    return out


def rel_pos_2d(table_v, table_h, n_tokens: int, max_rel: int, head_dim: int):
    """RelativePosition2D_super.forward, multihead_super.py:40-66: (N, N, head_dim)."""
    grid = int(math.isqrt(n_tokens - 1))
    iv, ih = rel_index.autoformer_rel_index(grid, max_rel)
    iv = torch.from_numpy(iv)
    ih = torch.from_numpy(ih)
    return table_v[:, :head_dim][iv] + table_h[:, :head_dim][ih]


def attention_core_autoformer(qkv5, tables, spec_max_rel: int, scale: float):
    """Attention core of AttentionSuper.forward, multihead_super.py:135-154, on
    qkv5 = qkv(x).reshape(B, N, 3, heads, hd).  tables = None or (k_v, k_h, v_v, v_h)."""
    B, N, _, heads, hd = qkv5.shape
    qkv = qkv5.permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * scale
    if tables is not None:
        r_p_k = rel_pos_2d(tables[0], tables[1], N, spec_max_rel, hd)
        attn = attn + (q.permute(2, 0, 1, 3).reshape(N, heads * B, -1) @ r_p_k.transpose(2, 1)) \
            .transpose(1, 0).reshape(B, heads, N, N) * scale
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    if tables is not None:
        r_p_v = rel_pos_2d(tables[2], tables[3], N, spec_max_rel, hd)
        attn_1 = attn.permute(2, 0, 1, 3).reshape(N, B * heads, -1)
        out = out + (attn_1 @ r_p_v).transpose(1, 0).reshape(B, heads, N, -1).transpose(2, 1).reshape(B, N, -1)
    return out


def attention_super(x, sd, prefix: str, spec: SupernetSpec, E: int, heads: int):
    """AttentionSuper.forward with change_qkv=True, multihead_super.py:100-116, 133-160;
    QKV slice per qkv_super.py:45-55, 72-83."""
    B, N, _ = x.shape
    qd = spec.head_dim * heads
    w = sd[prefix + "qkv.weight"][:, :E]
    w_s = torch.cat([w[i:3 * qd:3, :] for i in range(3)], dim=0)  # rows {i, i+3, ...} < 3*qd
    b_s = sd[prefix + "qkv.bias"][:3 * qd] if spec.qkv_bias else None  # contiguous, NOT interleaved
    qkv5 = F.linear(x, w_s, b_s).reshape(B, N, 3, heads, -1)
    scale = (qd // heads) ** -0.5
    tables = None
    if spec.relative_position:
        tables = tuple(sd[prefix + f"rel_pos_embed_{kv}.embeddings_table_{vh}"] for kv in "kv" for vh in "vh")
    out = attention_core_autoformer(qkv5, tables, spec.max_relative_position, scale)
    pw = sd[prefix + "proj.weight"][:E, :qd]   # Linear_super.py:71-75 top-left view
    pb = sd[prefix + "proj.bias"][:E]
    return F.linear(out, pw, pb)


def supernet_forward(sd: dict, config: dict, images: torch.Tensor, spec: SupernetSpec) -> torch.Tensor:
    """Vision_TransformerSuper.forward after set_sample_config(config)
    (supernet_transformer.py:102-127, 147-172, 225-287), drop / drop_path = 0."""
    E0 = config["embed_dim"][0]
    B = images.shape[0]
    x = F.conv2d(images, sd["patch_embed_super.proj.weight"][:E0], sd["patch_embed_super.proj.bias"][:E0],
                 stride=spec.patch_size).flatten(2).transpose(1, 2)       # embedding_super.py:33-40
    cls = sd["cls_token"][..., :E0].expand(B, -1, -1)
    x = torch.cat((cls, x), dim=1)
    if spec.abs_pos:
        x = x + sd["pos_embed"][..., :E0]
    out_dims = list(config["embed_dim"][1:]) + [config["embed_dim"][-1]]
    for i in range(config["layer_num"]):                                  # blocks >= layer_num are identity
        p = f"blocks.{i}."
        E = config["embed_dim"][i]
        heads = config["num_heads"][i]
        ffn = int(E * config["mlp_ratio"][i])
        res = x
        h = F.layer_norm(x, (E,), sd[p + "attn_layer_norm.weight"][:E], sd[p + "attn_layer_norm.bias"][:E], 1e-5)
        x = res + attention_super(h, sd, p + "attn.", spec, E, heads)
        res = x
        h = F.layer_norm(x, (E,), sd[p + "ffn_layer_norm.weight"][:E], sd[p + "ffn_layer_norm.bias"][:E], 1e-5)
        h = F.gelu(F.linear(h, sd[p + "fc1.weight"][:ffn, :E], sd[p + "fc1.bias"][:ffn]))
        h = F.linear(h, sd[p + "fc2.weight"][:out_dims[i], :ffn], sd[p + "fc2.bias"][:out_dims[i]])
        x = res + h
    El = config["embed_dim"][-1]
    x = F.layer_norm(x, (El,), sd["norm.weight"][:El], sd["norm.bias"][:El], 1e-5)
    x = torch.mean(x[:, 1:], dim=1) if spec.gp else x[:, 0]
    return F.linear(x, sd["head.weight"][:spec.num_classes, :El], sd["head.bias"][:spec.num_classes])


def sampled_param_count(config: dict, spec: SupernetSpec) -> int:
    """Vision_TransformerSuper.get_sampled_params_numel, supernet_transformer.py:129-138
    (sums calc_sampled_param_num of sampled layers + embed*(2+num_patches))."""
    E0 = config["embed_dim"][0]
    n = E0 * spec.in_chans * spec.patch_size ** 2 + E0
    for i in range(config["layer_num"]):
        E = config["embed_dim"][i]
        qd = spec.head_dim * config["num_heads"][i]
        ffn = int(E * config["mlp_ratio"][i])
        out = (list(config["embed_dim"][1:]) + [config["embed_dim"][-1]])[i]
        n += 3 * qd * E + 3 * qd            # qkv
        n += E * qd + E                      # proj
        n += 4 * (2 * spec.max_relative_position + 2) * (qd // config["num_heads"][i]) if spec.relative_position else 0
        n += 4 * E                           # two LayerNorms
        n += ffn * E + ffn + out * ffn + out
    El = config["embed_dim"][-1]
    n += 2 * El + spec.num_classes * El + spec.num_classes
    return n + E0 * (2 + spec.grid * spec.grid)


# ------------------------------------------------------------------------------------------
# DeiT + iRPE attention
# ------------------------------------------------------------------------------------------
@dataclass
class IrpeSpec:
    """get_rpe_config(...) for one of q/k/v, irpe.py:770-887."""
    ratio: float = 1.9
    method: int = rel_index.PRODUCT
    mode: str = "contextual"     # or "bias"
    shared_head: bool = True
    skip: int = 1
    rpe_on: str = "k"

    @property
    def alpha(self): return 1 * self.ratio
    @property
    def beta(self): return 2 * self.ratio
    @property
    def gamma(self): return 8 * self.ratio

    def num_buckets(self) -> int:
        return rel_index.irpe_num_buckets(self.method, self.beta) + (1 if self.skip > 0 else 0)


def irpe_rpe_transposed(x, table, bucket_ids, mode: str):
    """iRPE.forward_rpe_transpose, irpe.py:585-647.  x (B,H,L,D); contextual table
    (H or 1, D, nb); bias table (H or 1, nb)."""
    B, H, L, D = x.shape
    ids = torch.as_tensor(bucket_ids, dtype=torch.long).to(x.device)
    if mode == "bias":
        return table[:, ids.flatten()].view(1, table.shape[0], L, L)
    lookup = torch.matmul(x.transpose(0, 1).reshape(-1, B * L, D), table).view(-1, B, L, table.shape[-1]).transpose(0, 1)
    return torch.gather(lookup, 3, ids[None, None].expand(B, lookup.shape[1], L, L))   # = RPEIndexFunction


def irpe_rpe_value(attn, table, bucket_ids):
    """iRPE.forward_rpe_no_transpose, irpe.py:649-687.  table (H or 1, nb, D)."""
    ids = torch.as_tensor(bucket_ids, dtype=torch.long).to(table.device)
    L = ids.shape[0]
    weight = table[:, ids.flatten()].view(table.shape[0], L, L, table.shape[-1])
    return torch.matmul(attn.permute(1, 2, 0, 3), weight).permute(2, 0, 1, 3)


def rpe_attention(x, qkv_w, qkv_b, proj_w, proj_b, num_heads: int, bucket_ids,
                  rpe_q=None, rpe_k=None, rpe_v=None, mode: str = "contextual", return_core: bool = False):
    """RPEAttention.forward, iRPE/DeiT-with-iRPE/rpe_vision_transformer.py:68-97."""
    B, N, C = x.shape
    qkv = F.linear(x, qkv_w, qkv_b).reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    scale = (C // num_heads) ** -0.5
    q = q * scale
    attn = q @ k.transpose(-2, -1)
    # cross method (iRPE_Cross.forward, irpe.py:726-751): every table is a (rows, cols) pair with
    # its own bucket ids and the two encodings are summed
    def both(fn, x_, tab):
        if isinstance(tab, (tuple, list)):
            return fn(x_, tab[0], bucket_ids[0]) + fn(x_, tab[1], bucket_ids[1])
        return fn(x_, tab, bucket_ids)
    if rpe_k is not None:
        attn = attn + both(lambda x_, t, i: irpe_rpe_transposed(x_, t, i, mode), q, rpe_k)
    if rpe_q is not None:
        attn = attn + both(lambda x_, t, i: irpe_rpe_transposed(x_, t, i, mode), k * scale, rpe_q).transpose(2, 3)
    attn = attn.softmax(dim=-1)
    out = attn @ v
    if rpe_v is not None:
        out = out + both(irpe_rpe_value, attn, rpe_v)
    if return_core:
        return out
    x = out.transpose(1, 2).reshape(B, N, C)
    return F.linear(x, proj_w, proj_b)


def clip_attention(x, in_proj_weight, in_proj_bias, out_w, out_b, n_head: int, attn_mask=None,
                   head_z=None, hidden_z=None):
    """TinyCLIP ResidualAttentionBlock.attention, TinyCLIP/src/open_clip/model.py:238-283 (the naive
    branch, which also documents what the nn.MultiheadAttention branch computes).  x is
    (length, batch, embed) — the reference's LND layout; attn_mask an additive (length, length)
    float mask (text tower: -inf above the diagonal, model.py:756-762)."""
    length, batch, d_model = x.shape
    ws, bs = in_proj_weight.chunk(3), in_proj_bias.chunk(3)
    dph = ws[0].shape[0] // n_head
    q, k, v = [F.linear(x, w, b) for w, b in zip(ws, bs)]
    q, k, v = [t.reshape(length, batch * n_head, -1).transpose(0, 1) for t in (q, k, v)]
    sim = (q * dph ** -0.5) @ k.transpose(1, 2)
    if attn_mask is not None:
        sim = sim + attn_mask
    out = torch.softmax(sim, -1) @ v
    if head_z is not None:
        out = (out.view(batch, n_head, length, dph) * head_z.view(1, -1, 1, 1)).view(batch * n_head, length, dph)
    out = F.linear(out.transpose(0, 1).reshape(length, batch, -1), out_w, out_b)
    if hidden_z is not None:
        out = out * hidden_z
    return out


def tinyvit_bias_idxs(resolution):
    """attention_bias_idxs of TinyViT Attention.__init__, TinyViT/models/tiny_vit.py:237-252: offsets
    (|dr|, |dc|) numbered in first-seen order over the row-major point pairs; returns (idxs (N, N)
    int64, number of distinct offsets)."""
    import itertools
    points = list(itertools.product(range(resolution[0]), range(resolution[1])))
    offsets, idxs = {}, []
    for p1 in points:
        for p2 in points:
            off = (abs(p1[0] - p2[0]), abs(p1[1] - p2[1]))
            if off not in offsets:
                offsets[off] = len(offsets)
            idxs.append(offsets[off])
    n = len(points)
    return torch.tensor(idxs, dtype=torch.long).view(n, n), len(offsets)


def tinyvit_attention(x, P, num_heads: int, key_dim: int, d: int, idxs, eps: float = 1e-5):
    """TinyViT Attention.forward, TinyViT/models/tiny_vit.py:262-286.  P: norm.{weight,bias},
    qkv.{weight,bias}, proj.{weight,bias}, attention_biases (heads, n_offsets)."""
    B, N, C = x.shape
    x = F.layer_norm(x, (C,), P["norm.weight"], P["norm.bias"], eps)
    qkv = F.linear(x, P["qkv.weight"], P["qkv.bias"])
    q, k, v = qkv.view(B, N, num_heads, -1).split([key_dim, key_dim, d], dim=3)
    q, k, v = q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)
    attn = (q @ k.transpose(-2, -1)) * key_dim ** -0.5 + P["attention_biases"][:, idxs]
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, num_heads * d)
    return F.linear(x, P["proj.weight"], P["proj.bias"])
