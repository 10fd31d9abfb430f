"""Host-side mirror of AutoFormer/model/supernet_transformer.py for the B200 engine.

Same public surface as the reference (class names, constructor arguments, parameter names
and full-supernet shapes, set_sample_config / get_sampled_params_numel / get_complexity /
forward), so checkpoints, optimizers, DDP and supernet_engine.train_one_epoch work
unchanged.  Two execution paths over the SAME parameters:

  * fused=True (default): one autograd node for the whole sampled subnet, executed by
    cream_b200.engine (GEMM epilogues carry bias/GELU/DropPath/residual);
  * fused=False: module-by-module composition exactly like the reference's forward, each
    module being a drop-in from cream_b200.autoformer.model.module (this is also what
    runs when the reference's own supernet_transformer.py is imported on top of them).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import engine
from .module.Linear_super import LinearSuper
from .module.embedding_super import PatchembedSuper
from .module.layernorm_super import LayerNormSuper
from .module.multihead_super import AttentionSuper
from .utils import DropPath, drop_path_scale, trunc_normal_


def gelu(x: torch.Tensor) -> torch.Tensor:
    # fp32 GELU cast back to the input dtype (supernet_transformer.py:14-18)
    return F.gelu(x.float()).type_as(x)


def calc_dropout(dropout, sample_embed_dim, super_embed_dim):
    return dropout * 1.0 * sample_embed_dim / super_embed_dim


class TransformerEncoderLayer(nn.Module):
    """Pre/post-norm encoder block over sliceable modules (supernet_transformer.py:175-304)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, dropout=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, pre_norm=True, scale=False, relative_position=False,
                 change_qkv=False, max_relative_position=14):
        super().__init__()
        self.super_embed_dim = dim
        self.super_mlp_ratio = mlp_ratio
        self.super_ffn_embed_dim_this_layer = int(mlp_ratio * dim)
        self.super_num_heads = num_heads
        self.normalize_before = pre_norm
        self.super_dropout = attn_drop
        self.drop_path_prob = float(drop_path)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.scale = scale
        self.relative_position = relative_position
        self.sample_embed_dim = None
        self.sample_mlp_ratio = None
        self.sample_ffn_embed_dim_this_layer = None
        self.sample_num_heads_this_layer = None
        self.sample_scale = None
        self.sample_dropout = None
        self.sample_attn_dropout = None
        self.is_identity_layer = None
        self.attn = AttentionSuper(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                   attn_drop=attn_drop, proj_drop=dropout, scale=self.scale,
                                   relative_position=self.relative_position, change_qkv=change_qkv,
                                   max_relative_position=max_relative_position)
        self.attn_layer_norm = LayerNormSuper(self.super_embed_dim)
        self.ffn_layer_norm = LayerNormSuper(self.super_embed_dim)
        self.activation_fn = gelu
        self.fc1 = LinearSuper(super_in_dim=self.super_embed_dim, super_out_dim=self.super_ffn_embed_dim_this_layer)
        self.fc2 = LinearSuper(super_in_dim=self.super_ffn_embed_dim_this_layer, super_out_dim=self.super_embed_dim)

    def set_sample_config(self, is_identity_layer, sample_embed_dim=None, sample_mlp_ratio=None,
                          sample_num_heads=None, sample_dropout=None, sample_attn_dropout=None, sample_out_dim=None):
        if is_identity_layer:
            self.is_identity_layer = True
            return
        self.is_identity_layer = False
        self.sample_embed_dim = sample_embed_dim
        self.sample_out_dim = sample_out_dim
        self.sample_mlp_ratio = sample_mlp_ratio
        self.sample_ffn_embed_dim_this_layer = int(sample_embed_dim * sample_mlp_ratio)
        self.sample_num_heads_this_layer = sample_num_heads
        self.sample_dropout = sample_dropout
        self.sample_attn_dropout = sample_attn_dropout
        self.attn_layer_norm.set_sample_config(sample_embed_dim=self.sample_embed_dim)
        # 64 channels per head under change_qkv (supernet_transformer.py:243)
        self.attn.set_sample_config(sample_q_embed_dim=self.sample_num_heads_this_layer * 64,
                                    sample_num_heads=self.sample_num_heads_this_layer,
                                    sample_in_embed_dim=self.sample_embed_dim)
        self.fc1.set_sample_config(sample_in_dim=self.sample_embed_dim,
                                   sample_out_dim=self.sample_ffn_embed_dim_this_layer)
        self.fc2.set_sample_config(sample_in_dim=self.sample_ffn_embed_dim_this_layer,
                                   sample_out_dim=self.sample_out_dim)
        self.ffn_layer_norm.set_sample_config(sample_embed_dim=self.sample_embed_dim)

    def maybe_layer_norm(self, layer_norm, x, before=False, after=False):
        assert before ^ after
        return layer_norm(x) if (after ^ self.normalize_before) else x

    def forward(self, x):
        if self.is_identity_layer:
            return x
        residual = x
        x = self.maybe_layer_norm(self.attn_layer_norm, x, before=True)
        x = self.attn(x)
        x = F.dropout(x, p=self.sample_attn_dropout, training=self.training)
        x = self.drop_path(x)
        x = residual + x
        x = self.maybe_layer_norm(self.attn_layer_norm, x, after=True)
        residual = x
        x = self.maybe_layer_norm(self.ffn_layer_norm, x, before=True)
        x = self.activation_fn(self.fc1(x))
        x = F.dropout(x, p=self.sample_dropout, training=self.training)
        x = self.fc2(x)
        x = F.dropout(x, p=self.sample_dropout, training=self.training)
        if self.scale:
            x = x * (self.super_mlp_ratio / self.sample_mlp_ratio)
        x = self.drop_path(x)
        x = residual + x
        x = self.maybe_layer_norm(self.ffn_layer_norm, x, after=True)
        return x

    def get_complexity(self, sequence_length):
        if self.is_identity_layer:
            return 0
        total = self.attn_layer_norm.get_complexity(sequence_length + 1)
        total += self.attn.get_complexity(sequence_length + 1)
        total += self.ffn_layer_norm.get_complexity(sequence_length + 1)
        total += self.fc1.get_complexity(sequence_length + 1)
        total += self.fc2.get_complexity(sequence_length + 1)
        return total


class Vision_TransformerSuper(nn.Module):
    """Weight-entangled ViT supernet (supernet_transformer.py:21-172) on the B200 engine."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., pre_norm=True, scale=False, gp=False, relative_position=False,
                 change_qkv=False, abs_pos=True, max_relative_position=14, fused=True):
        super().__init__()
        self.super_embed_dim = embed_dim
        self.super_mlp_ratio = mlp_ratio
        self.super_layer_num = depth
        self.super_num_heads = num_heads
        self.super_dropout = drop_rate
        self.super_attn_dropout = attn_drop_rate
        self.num_classes = num_classes
        self.pre_norm = pre_norm
        self.scale = scale
        self.patch_embed_super = PatchembedSuper(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                                 embed_dim=embed_dim)
        self.gp = gp
        self.fused = fused
        self.change_qkv = change_qkv
        self.relative_position = relative_position
        self.sample_embed_dim = None
        self.sample_mlp_ratio = None
        self.sample_layer_num = None
        self.sample_num_heads = None
        self.sample_dropout = None
        self.sample_output_dim = None
        self.sample_config = None

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]   # stochastic depth decay rule
        self.blocks = nn.ModuleList([
            TransformerEncoderLayer(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, dropout=drop_rate, attn_drop=attn_drop_rate,
                                    drop_path=dpr[i], pre_norm=pre_norm, scale=self.scale, change_qkv=change_qkv,
                                    relative_position=relative_position,
                                    max_relative_position=max_relative_position)
            for i in range(depth)])
        num_patches = self.patch_embed_super.num_patches
        self.abs_pos = abs_pos
        if self.abs_pos:
            self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
            trunc_normal_(self.pos_embed, std=.02)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        trunc_normal_(self.cls_token, std=.02)
        if self.pre_norm:
            self.norm = LayerNormSuper(super_embed_dim=embed_dim)
        self.head = LinearSuper(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)

        self._geo = engine.SupernetGeometry(
            embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio,
            img_size=img_size if isinstance(img_size, int) else img_size[0],
            patch_size=patch_size if isinstance(patch_size, int) else patch_size[0], in_chans=in_chans,
            num_classes=num_classes, max_relative_position=max_relative_position, gp=gp,
            relative_position=relative_position, abs_pos=abs_pos, eps=self.blocks[0].attn_layer_norm.eps)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token', 'rel_pos_embed'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = LinearSuper(self.super_embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def set_sample_config(self, config: dict):
        self.sample_config = config
        self.sample_embed_dim = config['embed_dim']
        self.sample_mlp_ratio = config['mlp_ratio']
        self.sample_layer_num = config['layer_num']
        self.sample_num_heads = config['num_heads']
        self.sample_dropout = calc_dropout(self.super_dropout, self.sample_embed_dim[0], self.super_embed_dim)
        self.patch_embed_super.set_sample_config(self.sample_embed_dim[0])
        self.sample_output_dim = [d for d in self.sample_embed_dim[1:]] + [self.sample_embed_dim[-1]]
        for i, blk in enumerate(self.blocks):
            if i < self.sample_layer_num:
                blk.set_sample_config(
                    is_identity_layer=False, sample_embed_dim=self.sample_embed_dim[i],
                    sample_mlp_ratio=self.sample_mlp_ratio[i], sample_num_heads=self.sample_num_heads[i],
                    sample_dropout=calc_dropout(self.super_dropout, self.sample_embed_dim[i], self.super_embed_dim),
                    sample_out_dim=self.sample_output_dim[i],
                    sample_attn_dropout=calc_dropout(self.super_attn_dropout, self.sample_embed_dim[i],
                                                     self.super_embed_dim))
            else:
                blk.set_sample_config(is_identity_layer=True)
        if self.pre_norm:
            self.norm.set_sample_config(self.sample_embed_dim[-1])
        self.head.set_sample_config(self.sample_embed_dim[-1], self.num_classes)

    def get_sampled_params_numel(self, config):
        self.set_sample_config(config)
        numels = []
        for name, module in self.named_modules():
            if hasattr(module, 'calc_sampled_param_num'):
                parts = name.split('.')
                if parts[0] == 'blocks' and int(parts[1]) >= config['layer_num']:
                    continue
                numels.append(module.calc_sampled_param_num())
        return sum(numels) + self.sample_embed_dim[0] * (2 + self.patch_embed_super.num_patches)

    def get_complexity(self, sequence_length):
        total = self.patch_embed_super.get_complexity(sequence_length)
        total += np.prod(self.pos_embed[..., :self.sample_embed_dim[0]].size()) / 2.0
        for blk in self.blocks:
            total += blk.get_complexity(sequence_length + 1)
        total += self.head.get_complexity(sequence_length + 1)
        return total

    # ------------------------------------------------------------------ module-by-module path
    def forward_features(self, x):
        B = x.shape[0]
        x = self.patch_embed_super(x)
        cls_tokens = self.cls_token[..., :self.sample_embed_dim[0]].expand(B, -1, -1)
        x = torch.cat((cls_tokens, x.float()), dim=1)
        if self.abs_pos:
            x = x + self.pos_embed[..., :self.sample_embed_dim[0]]
        x = F.dropout(x, p=self.sample_dropout, training=self.training)
        for blk in self.blocks:
            x = blk(x)
        if self.pre_norm:
            x = self.norm(x)
        if self.gp:
            return torch.mean(x[:, 1:], dim=1)
        return x[:, 0]

    # ------------------------------------------------------------------ fused engine path
    def _fusable(self) -> bool:
        return (self.fused and self.pre_norm and self.change_qkv and not self.scale and self.num_classes > 0
                and self.super_dropout == 0.0 and self.super_attn_dropout == 0.0)

    def _drop_path_scales(self, batch, device):
        if not self.training or all(b.drop_path_prob == 0.0 for b in self.blocks):
            return None
        # floor(keep + U[0,1)) / keep per layer, branch and sample (model/utils.py:71-87), drawn for all
        # sampled layers in one shot: 4 small launches per step instead of 9 per layer
        blocks = self.blocks[:self.sample_layer_num]
        key = (tuple(b.drop_path_prob for b in blocks), str(device))
        cache = self.__dict__.setdefault("_keep_cache", {})
        keep = cache.get(key)
        if keep is None:
            keep = torch.tensor([1.0 - q for q in key[0]], dtype=torch.float32).view(-1, 1, 1).to(device)
            cache[key] = keep
        all_scales = torch.floor(keep + torch.rand(len(blocks), 2, batch, dtype=torch.float32, device=device)) / keep
        return [None if b.drop_path_prob == 0.0 else all_scales[i] for i, b in enumerate(blocks)]

    def forward(self, x):
        assert self.sample_config is not None, "call set_sample_config(config) first"
        if not self._fusable():
            return self.head(self.forward_features(x))
        if not x.is_cuda:
            raise RuntimeError("cream_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        P = dict(self.named_parameters())
        scales = self._drop_path_scales(x.shape[0], x.device)
        return engine.supernet_apply(P, self._geo, self.sample_config, x.float().contiguous(), scales)
