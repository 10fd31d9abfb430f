"""Fused-engine front end for the AutoFormer supernet.

There are two ways to put the B200 engine under the reference's training scripts:

1. `fuse_reference(cls_or_model)` — take the reference's OWN `Vision_TransformerSuper`
   (AutoFormer/model/supernet_transformer.py, imported unchanged, over either its own `model.module.*`
   or cream_b200's drop-ins) and route its `forward` through `cream_b200.engine` whenever the
   constructor flags describe something the fused kernels implement.  Nothing else of the class is
   touched: `set_sample_config`, `get_sampled_params_numel`, checkpoints, `no_weight_decay`, DDP.

2. `Vision_TransformerSuper` below — a stand-alone model with the same constructor, parameter names /
   full-supernet shapes (the checkpoint + optimizer contract) and the same public
   methods, for use where the reference checkout is not on the path (bench.py, the GPU box).  It is a
   parameter container plus a *plan*: `set_sample_config` turns the sampled config into one
   `LayerPlan` per block, and everything else (module configuration, counters, forward) reads the plan.

Both share `FusedSupernet`, which is the only place the engine is called from.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import engine
from .module.Linear_super import LinearSuper
from .module.embedding_super import PatchembedSuper
from .module.layernorm_super import LayerNormSuper
from .module.multihead_super import AttentionSuper
from .utils import DropPath, trunc_normal_

HEAD_DIM = 64   # the supernets pin 64 channels per head under change_qkv (supernet_transformer.py:243)


# ------------------------------------------------------------------------------------------------
# fused forward, shared by the mirror and by the patched reference class
# ------------------------------------------------------------------------------------------------
class FusedSupernet:
    """Mixin over an object with the reference's attribute names (`super_embed_dim`, `blocks`,
    `patch_embed_super`, `gp`, `abs_pos`, `relative_position`, `change_qkv`, `pre_norm`, `scale`,
    `super_dropout`, `super_attn_dropout`, `num_classes`, `sample_*`)."""

    fused = True

    def engine_geometry(self) -> engine.SupernetGeometry:
        pe, blk = self.patch_embed_super, self.blocks[0]
        one = lambda v: v if isinstance(v, int) else v[0]
        geo = engine.SupernetGeometry(
            embed_dim=self.super_embed_dim, depth=len(self.blocks), num_heads=self.super_num_heads,
            mlp_ratio=self.super_mlp_ratio, img_size=one(pe.img_size), patch_size=one(pe.patch_size),
            in_chans=pe.proj.in_channels, num_classes=self.num_classes,
            max_relative_position=getattr(blk.attn, "max_relative_position", 14), gp=bool(self.gp),
            relative_position=bool(getattr(blk.attn, "relative_position", False)), abs_pos=bool(self.abs_pos),
            eps=blk.attn_layer_norm.eps, qkv_bias=blk.attn.qkv.bias is not None)
        return geo

    def fusable(self) -> bool:
        """True when the constructor flags describe what the fused engine computes: pre-norm blocks,
        64 channels per head (change_qkv), no output rescaling, no dropout (DropPath is supported)."""
        blk = self.blocks[0]
        return bool(self.fused and getattr(self, "pre_norm", True) and getattr(blk.attn, "change_qkv", False)
                    and not getattr(self, "scale", False) and self.num_classes > 0
                    and self.super_dropout == 0.0 and self.super_attn_dropout == 0.0
                    and isinstance(self.head, nn.Linear))

    def drop_path_scales(self, batch: int, device) -> Optional[List[Optional[torch.Tensor]]]:
        """Per sampled layer a (2, batch) fp32 tensor floor(keep + U) / keep for the two residual
        branches (model/utils.py:71-87), drawn for all layers in one launch; None when inactive."""
        n = self.sample_layer_num
        probs = tuple(float(getattr(b.drop_path, "drop_prob", 0.0) or 0.0) for b in list(self.blocks)[:n])
        if not self.training or not any(probs):
            return None
        cache = self.__dict__.setdefault("_keep_cache", {})
        key = (probs, str(device))
        if key not in cache:
            cache[key] = torch.tensor([1.0 - q for q in probs], dtype=torch.float32).view(-1, 1, 1).to(device)
        keep = cache[key]
        draw = torch.floor(keep + torch.rand(n, 2, batch, dtype=torch.float32, device=device)) / keep
        return [None if q == 0.0 else draw[i] for i, q in enumerate(probs)]

    def sampled_config(self) -> dict:
        assert self.sample_layer_num is not None, "call set_sample_config(config) first"
        return {"layer_num": self.sample_layer_num, "embed_dim": list(self.sample_embed_dim),
                "num_heads": list(self.sample_num_heads), "mlp_ratio": list(self.sample_mlp_ratio)}

    def fused_forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("cream_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        geo = self.__dict__.get("_geo")
        if geo is None or geo.num_classes != self.num_classes:
            geo = self._geo = self.engine_geometry()
        params = dict(self.named_parameters())
        scales = self.drop_path_scales(x.shape[0], x.device)
        return engine.supernet_apply(params, geo, self.sampled_config(), x.float().contiguous(), scales, owner=self)


def fuse_reference(target):
    """Patch the reference's `Vision_TransformerSuper` (class or instance) so that `forward` runs the
    fused engine when `fusable()`, and its own module-by-module forward otherwise.  Returns `target`."""
    cls = target if isinstance(target, type) else type(target)
    if getattr(cls, "_cream_fused", False):
        return target
    stock_forward = cls.forward

    def forward(self, x):
        if FusedSupernet.fusable(self) and x.is_cuda:
            return FusedSupernet.fused_forward(self, x)
        return stock_forward(self, x)

    for name in ("engine_geometry", "fusable", "drop_path_scales", "sampled_config", "fused_forward"):
        setattr(cls, name, getattr(FusedSupernet, name))
    cls.fused, cls.forward, cls.stock_forward, cls._cream_fused = True, forward, stock_forward, True
    return target


# ------------------------------------------------------------------------------------------------
# stand-alone mirror
# ------------------------------------------------------------------------------------------------
@dataclass
class LayerPlan:
    """What one sampled block looks like (None for an identity block)."""
    embed: int
    out: int
    heads: int
    ratio: float
    dropout: float
    attn_dropout: float

    @property
    def ffn(self) -> int:
        return int(self.embed * self.ratio)

    @property
    def qk_width(self) -> int:
        return HEAD_DIM * self.heads


class SupernetBlock(nn.Module):
    """Parameters of one encoder block under the reference's names (`attn`, `attn_layer_norm`,
    `ffn_layer_norm`, `fc1`, `fc2`), configured from a LayerPlan."""

    def __init__(self, dim, heads, mlp_ratio, qkv_bias, qk_scale, dropout, attn_drop, drop_path, pre_norm, scale,
                 relative_position, change_qkv, max_relative_position):
        super().__init__()
        hidden = int(mlp_ratio * dim)
        self.super_mlp_ratio, self.normalize_before, self.scale = mlp_ratio, pre_norm, scale
        self.attn = AttentionSuper(dim, num_heads=heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                   proj_drop=dropout, scale=scale, relative_position=relative_position,
                                   change_qkv=change_qkv, max_relative_position=max_relative_position)
        self.attn_layer_norm, self.ffn_layer_norm = LayerNormSuper(dim), LayerNormSuper(dim)
        self.fc1, self.fc2 = LinearSuper(dim, hidden), LinearSuper(hidden, dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.plan: Optional[LayerPlan] = None
        self.is_identity_layer = None

    def configure(self, plan: Optional[LayerPlan]) -> None:
        self.plan, self.is_identity_layer = plan, plan is None
        if plan is None:
            return
        for ln in (self.attn_layer_norm, self.ffn_layer_norm):
            ln.set_sample_config(sample_embed_dim=plan.embed)
        self.attn.set_sample_config(sample_q_embed_dim=plan.qk_width, sample_num_heads=plan.heads,
                                    sample_in_embed_dim=plan.embed)
        self.fc1.set_sample_config(sample_in_dim=plan.embed, sample_out_dim=plan.ffn)
        self.fc2.set_sample_config(sample_in_dim=plan.ffn, sample_out_dim=plan.out)

    # module-by-module evaluation (used when the fused engine does not apply)
    def _mlp(self, x):
        p = self.plan
        h = F.gelu(self.fc1(x).float()).type_as(x)          # GELU in fp32, cast back (:14-18 of the reference)
        h = F.dropout(h, p=p.dropout, training=self.training)
        h = F.dropout(self.fc2(h), p=p.dropout, training=self.training)
        return h * (self.super_mlp_ratio / p.ratio) if self.scale else h

    def _attention(self, x):
        return F.dropout(self.attn(x), p=self.plan.attn_dropout, training=self.training)

    def forward(self, x):
        if self.plan is None:
            return x
        for norm, branch in ((self.attn_layer_norm, self._attention), (self.ffn_layer_norm, self._mlp)):
            y = branch(norm(x) if self.normalize_before else x)
            x = x + self.drop_path(y)
            if not self.normalize_before:
                x = norm(x)
        return x

    def flops(self, tokens: int) -> float:
        if self.plan is None:
            return 0
        parts = (self.attn_layer_norm, self.attn, self.ffn_layer_norm, self.fc1, self.fc2)
        return sum(m.get_complexity(tokens + 1) for m in parts)

    get_complexity = flops


class Vision_TransformerSuper(FusedSupernet, nn.Module):
    """Weight-entangled ViT supernet with the reference's constructor and state_dict
    (supernet_transformer.py:21-78), executed by the fused engine."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., pre_norm=True, scale=False, gp=False, relative_position=False,
                 change_qkv=False, abs_pos=True, max_relative_position=14, fused=True):
        super().__init__()
        self.super_embed_dim, self.super_mlp_ratio = embed_dim, mlp_ratio
        self.super_layer_num, self.super_num_heads = depth, num_heads
        self.super_dropout, self.super_attn_dropout = drop_rate, attn_drop_rate
        self.num_classes, self.pre_norm, self.scale, self.gp = num_classes, pre_norm, scale, gp
        self.fused, self.change_qkv, self.relative_position, self.abs_pos = fused, change_qkv, relative_position, abs_pos
        self.patch_embed_super = PatchembedSuper(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                                 embed_dim=embed_dim)
        rates = torch.linspace(0, drop_path_rate, depth).tolist()      # stochastic-depth decay rule
        self.blocks = nn.ModuleList(
            SupernetBlock(embed_dim, num_heads, mlp_ratio, qkv_bias, qk_scale, drop_rate, attn_drop_rate, rate,
                          pre_norm, scale, relative_position, change_qkv, max_relative_position) for rate in rates)
        tokens = self.patch_embed_super.num_patches + 1
        if abs_pos:
            self.pos_embed = nn.Parameter(trunc_normal_(torch.zeros(1, tokens, embed_dim), std=.02))
        self.cls_token = nn.Parameter(trunc_normal_(torch.zeros(1, 1, embed_dim), std=.02))
        if pre_norm:
            self.norm = LayerNormSuper(super_embed_dim=embed_dim)
        self.head = LinearSuper(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        for m in self.modules():      # Linear: trunc-normal .02 / zero bias; LayerNorm: ones / zeros
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        self.sample_config = None
        self.sample_embed_dim = self.sample_mlp_ratio = self.sample_num_heads = None
        self.sample_layer_num = self.sample_dropout = self.sample_output_dim = None

    # ---- reference surface -------------------------------------------------------------------
    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token', 'rel_pos_embed'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = LinearSuper(self.super_embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.__dict__.pop("_geo", None)
        if num_classes > 0:
            self.head.to(self.cls_token.device)

    def plan(self, config: dict) -> List[Optional[LayerPlan]]:
        """One LayerPlan per block; blocks past `layer_num` are identities.  A block's output width is
        the NEXT block's embed dim; dropout probabilities shrink with the sampled width."""
        dims, depth = config['embed_dim'], config['layer_num']
        shrink = lambda p, e: p * 1.0 * e / self.super_embed_dim
        plans: List[Optional[LayerPlan]] = []
        for i in range(len(self.blocks)):
            if i >= depth:
                plans.append(None)
                continue
            nxt = dims[i + 1] if i + 1 < len(dims) else dims[-1]
            plans.append(LayerPlan(dims[i], nxt, config['num_heads'][i], config['mlp_ratio'][i],
                                   shrink(self.super_dropout, dims[i]), shrink(self.super_attn_dropout, dims[i])))
        return plans

    def set_sample_config(self, config: dict):
        self.sample_config = config
        self.sample_embed_dim, self.sample_mlp_ratio = config['embed_dim'], config['mlp_ratio']
        self.sample_layer_num, self.sample_num_heads = config['layer_num'], config['num_heads']
        dims = self.sample_embed_dim
        self.sample_dropout = self.super_dropout * 1.0 * dims[0] / self.super_embed_dim
        self.sample_output_dim = list(dims[1:]) + [dims[-1]]
        self.patch_embed_super.set_sample_config(dims[0])
        for block, plan in zip(self.blocks, self.plan(config)):
            block.configure(plan)
        if self.pre_norm:
            self.norm.set_sample_config(dims[-1])
        if self.num_classes > 0:
            self.head.set_sample_config(dims[-1], self.num_classes)

    def get_sampled_params_numel(self, config):
        """Element count of the sampled sub-network (what evolution.py:90 constrains): every sliced
        view of the sampled blocks, the embedding, the final norm and the head, plus the sampled
        columns of cls_token / pos_embed."""
        self.set_sample_config(config)
        live = [self.patch_embed_super, self.head] + ([self.norm] if self.pre_norm else [])
        for block in list(self.blocks)[:config['layer_num']]:
            live += [m for m in block.modules() if m is not block]
        counted = sum(m.calc_sampled_param_num() for m in live if hasattr(m, 'calc_sampled_param_num'))
        return counted + config['embed_dim'][0] * (2 + self.patch_embed_super.num_patches)

    def get_complexity(self, sequence_length):
        pos = self.pos_embed[..., :self.sample_embed_dim[0]].numel() / 2.0
        body = sum(block.flops(sequence_length + 1) for block in self.blocks)
        return (self.patch_embed_super.get_complexity(sequence_length) + pos + body
                + self.head.get_complexity(sequence_length + 1))

    # ---- evaluation --------------------------------------------------------------------------
    def forward_features(self, x):
        """Module-by-module path over the drop-in modules (each a cream_b200 kernel)."""
        e0 = self.sample_embed_dim[0]
        tokens = self.patch_embed_super(x)
        cls = self.cls_token[..., :e0].expand(tokens.shape[0], -1, -1)
        x = torch.cat((cls, tokens.float()), dim=1)
        if self.abs_pos:
            x = x + self.pos_embed[..., :e0]
        x = F.dropout(x, p=self.sample_dropout, training=self.training)
        for block in self.blocks:
            x = block(x)
        if self.pre_norm:
            x = self.norm(x)
        return x[:, 1:].mean(dim=1) if self.gp else x[:, 0]

    def forward(self, x):
        assert self.sample_config is not None, "call set_sample_config(config) first"
        if self.fusable():
            return self.fused_forward(x)
        if not x.is_cuda:
            raise RuntimeError("cream_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        return self.head(self.forward_features(x))
