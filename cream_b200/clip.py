"""TinyCLIP ViT towers and contrastive step on the fused B200 kernels.

The classes carry the reference's parameter names (TinyCLIP/src/open_clip/model.py):
`VisualTransformer` (model.py:442-536: conv1, class_embedding, positional_embedding, ln_pre,
transformer.resblocks.N.{ln_1, attn.in_proj_*, attn.out_proj, ln_2, mlp.c_fc, mlp.c_proj}, ln_post,
proj), `TextEncoder` (model.py:682-806: token_embedding, positional_embedding, transformer, ln_final,
text_projection, causal mask), `CLIP` (model.py:874-1001,1073-1112: `_image_encoder.visual`,
`_text_encoder`, `_logit_scale.logit_scale`; forward returns normalised features and exp(logit_scale)),
`ClipLoss` / `gather_features` (loss.py:18-66,110-166).

Each tower is ONE autograd node: its forward and backward are sequences of cream_b200 launches
(im2col + GEMM for the 32x32 stride-32 convolution, LayerNorm, QKV GEMM, the fused attention kernel
with the text tower's causal mask computed from the token coordinates, projection / MLP GEMMs with fused bias, GELU and fp32
residual epilogues).  The residual stream, LayerNorm statistics and weight gradients are fp32, GEMM
operands bf16 - the arithmetic of the reference under `--precision amp`.  torch carries the token
embedding lookup, the eot-row selection and the (B x B) contrastive logits.

Not supported (asserted): QuickGELU, the pruning masks (`hidden_z` ... `embed_dim_z`), timm / ResNet
image towers.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.distributed.nn
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from ._lib import EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_F32, EPI_F32_RESID, check

_p, _stream = ops._p, ops._stream
HD = ops.HEAD_DIM


def block_names(prefix: str, i: int) -> List[str]:
    p = f"{prefix}resblocks.{i}."
    return [p + "ln_1.weight", p + "ln_1.bias", p + "attn.in_proj_weight", p + "attn.in_proj_bias",
            p + "attn.out_proj.weight", p + "attn.out_proj.bias", p + "ln_2.weight", p + "ln_2.bias",
            p + "mlp.c_fc.weight", p + "mlp.c_fc.bias", p + "mlp.c_proj.weight", p + "mlp.c_proj.bias"]


# ------------------------------------------------------------------------------------------------
# transformer blocks (model.py:286-328: x += attn(ln_1(x)); x += mlp(ln_2(x)))
# ------------------------------------------------------------------------------------------------
def _attn_packing(B, N, causal):
    """(batch, tokens, block_len) handed to the attention kernel.  Its query tile is 128 rows: with the image
    tower's 50 tokens a tile would be 39 % full, so two batch items are passed as ONE 100-token sequence with
    block-diagonal visibility (cream_attn_desc.block_len) - the token-major (B*N, .) buffers are unchanged.
    The 77-token text tower does not fit twice into a tile and stays one item per tile."""
    if B % 2 == 0 and 2 * N <= 128:
        return B // 2, 2 * N, N
    return B, N, 0


def blocks_forward(P, prefix, layers, heads, x, B, N, causal, save):
    """x: (B*N, E) fp32 residual stream.  Returns (x_out, saved-per-block list)."""
    M, E = x.shape
    sh = ops.SHADOWS
    ffn = P[f"{prefix}resblocks.0.mlp.c_fc.weight"].shape[0]
    eps = 1e-5
    scale = HD ** -0.5
    saved = []
    Ba, Na, blk = _attn_packing(B, N, causal)
    for i in range(layers):
        p = f"{prefix}resblocks.{i}."
        ln1, mu1, rs1 = ops.layernorm_fwd(x, P[p + "ln_1.weight"], P[p + "ln_1.bias"], eps, E, save_stats=save)
        qkv = ops.linear_fwd(ln1, sh.get(P[p + "attn.in_proj_weight"]), 3 * E, E, P[p + "attn.in_proj_bias"])
        att, lse = ops.attention_fwd(qkv, Ba, heads, Na, scale, causal=causal, need_lse=save, block=blk)
        x1 = ops.linear_fwd(att, sh.get(P[p + "attn.out_proj.weight"]), E, E, P[p + "attn.out_proj.bias"],
                            epi=EPI_F32_RESID, resid=x)
        ln2, mu2, rs2 = ops.layernorm_fwd(x1, P[p + "ln_2.weight"], P[p + "ln_2.bias"], eps, E, save_stats=save)
        hpre = ops.empty_bf16(M, ffn, x.device)
        act = ops.linear_fwd(ln2, sh.get(P[p + "mlp.c_fc.weight"]), ffn, E, P[p + "mlp.c_fc.bias"],
                             epi=EPI_BF16_GELU, aux=hpre)
        x2 = ops.linear_fwd(act, sh.get(P[p + "mlp.c_proj.weight"]), E, ffn, P[p + "mlp.c_proj.bias"],
                            epi=EPI_F32_RESID, resid=x1)
        if save:
            saved.append(dict(x=x, ln1=ln1, mu1=mu1, rs1=rs1, qkv=qkv, att=att, lse=lse, x1=x1, ln2=ln2, mu2=mu2,
                              rs2=rs2, hpre=hpre, act=act))
        x = x2
    return x, saved


def blocks_backward(P, G, prefix, layers, heads, saved, g, B, N, causal):
    """g: (B*N, E) fp32 gradient of the block stack's output; returns the gradient of its input."""
    M, E = g.shape
    sh = ops.SHADOWS
    ffn = P[f"{prefix}resblocks.0.mlp.c_fc.weight"].shape[0]
    scale = HD ** -0.5
    Ba, Na, blk = _attn_packing(B, N, causal)
    dy2 = None
    for i in reversed(range(layers)):
        p = f"{prefix}resblocks.{i}."
        s = saved[i]
        if dy2 is None:     # top block; below it the bf16 copy comes out of the LayerNorm backward of the block above
            dy2 = ops.cast_scale(g, dbias=G[p + "mlp.c_proj.bias"])
        ops.linear_wgrad(dy2, s["act"], E, ffn, G[p + "mlp.c_proj.weight"])
        dh = ops.linear_dgrad(dy2, sh.get(P[p + "mlp.c_proj.weight"]), E, ffn, epi=EPI_BF16_DGELU, aux=s["hpre"])
        ops.bias_grad(dh, G[p + "mlp.c_fc.bias"])
        ops.linear_wgrad(dh, s["ln2"], ffn, E, G[p + "mlp.c_fc.weight"])
        dln2 = ops.linear_dgrad(dh, sh.get(P[p + "mlp.c_fc.weight"]), ffn, E)
        g1, dy1 = ops.layernorm_bwd_cast(dln2, s["x1"], P[p + "ln_2.weight"], s["mu2"], s["rs2"], E, G[p + "ln_2.weight"],
                                         G[p + "ln_2.bias"], resid_grad=g, dbias=G[p + "attn.out_proj.bias"])
        ops.linear_wgrad(dy1, s["att"], E, E, G[p + "attn.out_proj.weight"])
        datt = ops.linear_dgrad(dy1, sh.get(P[p + "attn.out_proj.weight"]), E, E)
        dqkv = ops.attention_bwd(s["qkv"], s["att"], s["lse"], datt, Ba, heads, Na, scale, causal=causal, block=blk)[0]
        ops.bias_grad(dqkv, G[p + "attn.in_proj_bias"])
        ops.linear_wgrad(dqkv, s["ln1"], 3 * E, E, G[p + "attn.in_proj_weight"])
        dln1 = ops.linear_dgrad(dqkv, sh.get(P[p + "attn.in_proj_weight"]), 3 * E, E)
        if i > 0:
            g, dy2 = ops.layernorm_bwd_cast(dln1, s["x"], P[p + "ln_1.weight"], s["mu1"], s["rs1"], E, G[p + "ln_1.weight"],
                                            G[p + "ln_1.bias"], resid_grad=g1,
                                            dbias=G[f"{prefix}resblocks.{i - 1}.mlp.c_proj.bias"])
        else:
            g = ops.layernorm_bwd(dln1, s["x"], P[p + "ln_1.weight"], s["mu1"], s["rs1"], E, G[p + "ln_1.weight"],
                                  G[p + "ln_1.bias"], resid_grad=g1)
        saved[i] = None
    return g


def _project(pooled_bf16, proj, out_dim):
    """pooled @ proj with proj stored (width, out_dim) as the reference keeps it (model.py:484,538-539)."""
    width = proj.shape[0]
    out = ops.empty_f32(pooled_bf16.shape[0], out_dim, pooled_bf16.device)
    ops.linear_dgrad(pooled_bf16, ops.SHADOWS.get(proj), width, out_dim, epi=EPI_F32, out=out)
    return out


def _project_backward(dfeat, pooled_bf16, proj, g_proj):
    df = ops.as_bf16_2d(dfeat)
    width, out_dim = proj.shape
    ops.linear_wgrad(pooled_bf16, df, width, out_dim, g_proj)          # d proj = pooled^T dfeat
    return ops.linear_fwd(df, ops.SHADOWS.get(proj), width, out_dim)     # d pooled = dfeat proj^T


def _grad_views(P: Dict[str, torch.Tensor], names) -> Dict[str, torch.Tensor]:
    """Zeroed fp32 gradients for `names`, carved out of ONE flat buffer (one memset instead of one per parameter);
    every view starts 16-byte aligned (the weight-gradient GEMMs reduce into them with bulk adds)."""
    sizes = [(P[n].numel() + 3) // 4 * 4 for n in names]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=P[names[0]].device)
    out, off = {}, 0
    for n, sz in zip(names, sizes):
        out[n] = flat[off:off + P[n].numel()].view(P[n].shape)
        off += sz
    return out


class _VisionTowerFn(torch.autograd.Function):
    """VisualTransformer.forward (model.py:493-536) without masks."""

    @staticmethod
    def forward(ctx, spec, names, images, *params):
        P = dict(zip(names, params))
        lib = _lib.load()
        dev = images.device
        B = images.shape[0]
        E, N, heads, layers, patch = spec["width"], spec["tokens"], spec["heads"], spec["layers"], spec["patch"]
        T = N - 1
        save = any(ctx.needs_input_grad)
        img = images.float().contiguous()
        kdim = img.shape[1] * patch * patch
        cols = ops.empty_bf16(B * T, kdim, dev)
        check(lib.cream_patch_im2col(_p(img), _p(cols), cols.stride(0), B, img.shape[1], img.shape[2], img.shape[3],
                                     patch, _stream()), "cream_patch_im2col")
        tok = ops.linear_fwd(cols, ops.SHADOWS.get(P["conv1.weight"]), E, kdim)
        x0 = ops.empty_f32(B * N, E, dev)
        check(lib.cream_tokens_assemble_fwd(_p(tok), tok.stride(0), _p(P["class_embedding"]), _p(P["positional_embedding"]),
                                            E, _p(x0), x0.stride(0), B, N, E, _stream()), "cream_tokens_assemble_fwd")
        x, mu0, rs0 = ops.layernorm_fwd(x0, P["ln_pre.weight"], P["ln_pre.bias"], 1e-5, E, out_f32=True, save_stats=save)
        x, blocks = blocks_forward(P, "transformer.", layers, heads, x, B, N, False, save)
        cls_rows = x.view(B, N, E)[:, 0]                                  # (B, E) view, pitch N*E
        pooled, mu_p, rs_p = ops.layernorm_fwd(cls_rows, P["ln_post.weight"], P["ln_post.bias"], 1e-5, E, save_stats=save)
        feat = _project(pooled, P["proj"], spec["out"])
        if save:
            ctx.spec, ctx.names, ctx.B = spec, names, B
            ctx.saved = dict(cols=cols, x0=x0, mu0=mu0, rs0=rs0, blocks=blocks, x_last=x, pooled=pooled, mu_p=mu_p,
                             rs_p=rs_p)
            ctx.save_for_backward(*params)
        return feat[:, :spec["out"]].contiguous()

    @staticmethod
    def backward(ctx, dfeat):
        P = dict(zip(ctx.names, ctx.saved_tensors))
        G = _grad_views(P, ctx.names)
        lib = _lib.load()
        spec, B, s = ctx.spec, ctx.B, ctx.saved
        E, N, heads, layers, patch = spec["width"], spec["tokens"], spec["heads"], spec["layers"], spec["patch"]
        dev = dfeat.device
        dpooled = _project_backward(dfeat.float().contiguous(), s["pooled"], P["proj"], G["proj"])
        cls_rows = s["x_last"].view(B, N, E)[:, 0]
        dcls = ops.layernorm_bwd(dpooled, cls_rows, P["ln_post.weight"], s["mu_p"], s["rs_p"], E, G["ln_post.weight"],
                                 G["ln_post.bias"])
        g = ops.empty_f32(B * N, E, dev, zero=True)
        g.view(B, N, E)[:, 0].copy_(dcls)
        g = blocks_backward(P, G, "transformer.", layers, heads, s["blocks"], g, B, N, False)
        g0 = ops.layernorm_bwd(g, s["x0"], P["ln_pre.weight"], s["mu0"], s["rs0"], E, G["ln_pre.weight"], G["ln_pre.bias"])
        dtok = ops.empty_bf16(B * (N - 1), E, dev)
        check(lib.cream_tokens_assemble_bwd(_p(g0), g0.stride(0), _p(dtok), dtok.stride(0), _p(G["positional_embedding"]), E,
                                            _p(G["class_embedding"]), B, N, E, _stream()), "cream_tokens_assemble_bwd")
        kdim = s["cols"].shape[1]
        ops.linear_wgrad(dtok, s["cols"], E, kdim, G["conv1.weight"])
        ctx.saved = None
        return (None, None, None) + tuple(G[n] for n in ctx.names)


class _TextTowerFn(torch.autograd.Function):
    """TextEncoder.encode_text (model.py:764-806) without masks."""

    @staticmethod
    def forward(ctx, spec, names, text, *params):
        P = dict(zip(names, params))
        dev = text.device
        B, N = text.shape
        E, heads, layers = spec["width"], spec["heads"], spec["layers"]
        save = any(ctx.needs_input_grad)
        x = ops.empty_f32(B * N, E, dev)
        x.view(B, N, E).copy_(P["token_embedding.weight"][text] + P["positional_embedding"][:N])
        x, blocks = blocks_forward(P, "transformer.", layers, heads, x, B, N, True, save)   # causal mask, model.py:756-762
        eot = text.argmax(dim=-1)                                          # the eot token has the largest id
        rows = torch.arange(B, device=dev) * N + eot
        x_eot = ops.empty_f32(B, E, dev)
        x_eot.copy_(x[rows])
        pooled, mu_p, rs_p = ops.layernorm_fwd(x_eot, P["ln_final.weight"], P["ln_final.bias"], 1e-5, E, save_stats=save)
        feat = _project(pooled, P["text_projection"], spec["out"])
        if save:
            ctx.spec, ctx.names = spec, names
            ctx.saved = dict(text=text, blocks=blocks, rows=rows, x_eot=x_eot, pooled=pooled, mu_p=mu_p, rs_p=rs_p)
            ctx.save_for_backward(*params)
        return feat[:, :spec["out"]].contiguous()

    @staticmethod
    def backward(ctx, dfeat):
        P = dict(zip(ctx.names, ctx.saved_tensors))
        G = _grad_views(P, ctx.names)
        spec, s = ctx.spec, ctx.saved
        text = s["text"]
        B, N = text.shape
        E, heads, layers = spec["width"], spec["heads"], spec["layers"]
        dev = dfeat.device
        dpooled = _project_backward(dfeat.float().contiguous(), s["pooled"], P["text_projection"], G["text_projection"])
        deot = ops.layernorm_bwd(dpooled, s["x_eot"], P["ln_final.weight"], s["mu_p"], s["rs_p"], E, G["ln_final.weight"],
                                 G["ln_final.bias"])
        g = ops.empty_f32(B * N, E, dev, zero=True)
        g[s["rows"]] = deot
        g = blocks_backward(P, G, "transformer.", layers, heads, s["blocks"], g, B, N, True)
        g3 = g.view(B, N, E)
        G["positional_embedding"][:N] = g3.sum(0)
        G["token_embedding.weight"].index_add_(0, text.reshape(-1), g)
        ctx.saved = None
        return (None, None, None) + tuple(G[n] for n in ctx.names)


# ------------------------------------------------------------------------------------------------
# modules with the reference's attribute / parameter names
# ------------------------------------------------------------------------------------------------
class LayerNorm(nn.LayerNorm):
    """Parameter holder (model.py:40-68); the towers run it inside their fused sequence."""


class _AttnParams(nn.Module):
    """The parameters of the nn.MultiheadAttention each block wraps (model.py:230)."""

    def __init__(self, width):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = nn.Linear(width, width)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class Mlp(nn.Module):
    def __init__(self, d_model, mlp_width):
        super().__init__()
        self.c_fc = nn.Linear(d_model, mlp_width)
        self.c_proj = nn.Linear(mlp_width, d_model)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, mlp_ratio=4.0):
        super().__init__()
        self.ln_1 = LayerNorm(d_model)
        self.attn = _AttnParams(d_model)
        self.ln_2 = LayerNorm(d_model)
        self.mlp = Mlp(d_model, int(d_model * mlp_ratio))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, mlp_ratio=4.0):
        super().__init__()
        assert width == heads * HD, "the fused attention kernel is built for head_dim 64"
        self.width, self.layers, self.num_heads, self.mlp_ratio = width, layers, heads, mlp_ratio
        self.head_dim = width // heads
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads, mlp_ratio) for _ in range(layers)])


def _tower_names(module: nn.Module) -> List[str]:
    return [n for n, _ in module.named_parameters()]


class VisualTransformer(nn.Module):
    def __init__(self, image_size, patch_size, width, layers, heads, mlp_ratio, output_dim, act_layer=nn.GELU):
        super().__init__()
        assert act_layer is nn.GELU, "QuickGELU is not part of the fused epilogue"
        self.image_size, self.patch_size = (image_size, image_size), (patch_size, patch_size)
        self.grid_size = (image_size // patch_size, image_size // patch_size)
        self.output_dim, self.embed_dim, self.layers = output_dim, width, layers
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid_size[0] * self.grid_size[1] + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads, mlp_ratio)
        self.head_dim = width // heads
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x, **masks):
        assert all(v is None for v in masks.values()), "pruning masks are not supported by the fused towers"
        if not x.is_cuda:
            raise RuntimeError("cream_b200 modules run on CUDA (sm_100a) tensors only")
        spec = dict(width=self.embed_dim, tokens=self.grid_size[0] * self.grid_size[1] + 1, heads=self.transformer.num_heads,
                    layers=self.layers, patch=self.patch_size[0], out=self.output_dim)
        names = _tower_names(self)
        P = dict(self.named_parameters())
        return _VisionTowerFn.apply(spec, names, x, *[P[n] for n in names])


class ImageEncoder(nn.Module):
    """model.py:597-678 for the ViT image tower."""

    def __init__(self, embed_dim, image_size=224, patch_size=32, width=768, layers=12, head_width=64, mlp_ratio=4.0):
        super().__init__()
        self.visual = VisualTransformer(image_size, patch_size, width, layers, width // head_width, mlp_ratio, embed_dim)

    def forward(self, image, normalized=False):
        f = self.visual(image)
        return F.normalize(f, dim=-1) if normalized else f


class TextEncoder(nn.Module):
    def __init__(self, embed_dim, context_length=77, vocab_size=49408, width=512, heads=8, layers=12):
        super().__init__()
        self.context_length, self.vocab_size = context_length, vocab_size
        self.transformer = Transformer(width, layers, heads)
        self.token_embedding = nn.Embedding(vocab_size, width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, width))
        self.ln_final = LayerNorm(width)
        self.text_projection = nn.Parameter(torch.empty(width, embed_dim))
        self.register_buffer("attn_mask", self.build_attention_mask(), persistent=False)
        self.init_parameters()

    def init_parameters(self):
        """model.py:738-755."""
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        t = self.transformer
        proj_std = (t.width ** -0.5) * ((2 * t.layers) ** -0.5)
        for b in t.resblocks:
            nn.init.normal_(b.attn.in_proj_weight, std=t.width ** -0.5)
            nn.init.normal_(b.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(b.mlp.c_fc.weight, std=(2 * t.width) ** -0.5)
            nn.init.normal_(b.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=t.width ** -0.5)

    def build_attention_mask(self):
        """Additive causal mask (model.py:756-762)."""
        return torch.full((self.context_length, self.context_length), float("-inf")).triu_(1)

    def encode_text(self, text, normalized=False):
        if not text.is_cuda:
            raise RuntimeError("cream_b200 modules run on CUDA (sm_100a) tensors only")
        t = self.transformer
        L = self.context_length
        spec = dict(width=t.width, heads=t.num_heads, layers=t.layers, out=self.text_projection.shape[1])
        names = _tower_names(self)
        P = dict(self.named_parameters())
        f = _TextTowerFn.apply(spec, names, text, *[P[n] for n in names])
        return F.normalize(f, dim=-1) if normalized else f

    def forward(self, text, normalized=False):
        return self.encode_text(text, normalized=normalized)


class LogitScale(nn.Module):
    def __init__(self):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    def forward(self, dummy=None):
        return self.logit_scale


class CLIP(nn.Module):
    """CLIP over the two fused towers; `vision_cfg` / `text_cfg` are the dicts of the reference's
    model_configs/*.json (e.g. ViT-B-32.json)."""

    def __init__(self, embed_dim: int, vision_cfg: dict, text_cfg: dict, quick_gelu: bool = False):
        super().__init__()
        assert not quick_gelu, "QuickGELU is not part of the fused epilogue"
        v = dict(image_size=224, layers=12, width=768, head_width=64, mlp_ratio=4.0, patch_size=16)
        v.update(vision_cfg)
        t = dict(context_length=77, vocab_size=49408, width=512, heads=8, layers=12)
        t.update(text_cfg)
        self._image_encoder = ImageEncoder(embed_dim, v["image_size"], v["patch_size"], v["width"], v["layers"],
                                           v["head_width"], v["mlp_ratio"])
        self._text_encoder = TextEncoder(embed_dim, t["context_length"], t["vocab_size"], t["width"], t["heads"], t["layers"])
        self._logit_scale = LogitScale()

    @property
    def visual(self):
        return self._image_encoder.visual

    @property
    def transformer(self):
        return self._text_encoder.transformer

    @property
    def logit_scale(self):
        return self._logit_scale.logit_scale

    def encode_image(self, image, normalized=False):
        return self._image_encoder(image, normalized=normalized)

    def encode_text(self, text, normalized=False):
        return self._text_encoder(text, normalized=normalized)

    def forward(self, image, text, normalized=True):
        image_features = self._image_encoder(image, normalized=normalized) if image is not None else None
        text_features = self._text_encoder(text, normalized=normalized) if text is not None else None
        return image_features, text_features, self._logit_scale().exp()


VIT_B_32 = dict(embed_dim=512, vision_cfg=dict(image_size=224, layers=12, width=768, patch_size=32),
                text_cfg=dict(context_length=77, vocab_size=49408, width=512, heads=8, layers=12))


# ------------------------------------------------------------------------------------------------
# contrastive loss with the cross-rank feature gather (loss.py)
# ------------------------------------------------------------------------------------------------
def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0, world_size=1):
    """loss.py:18-66 without the horovod branch."""
    if gather_with_grad:
        all_image = torch.cat(torch.distributed.nn.all_gather(image_features), dim=0)
        all_text = torch.cat(torch.distributed.nn.all_gather(text_features), dim=0)
    else:
        gi = [torch.zeros_like(image_features) for _ in range(world_size)]
        gt = [torch.zeros_like(text_features) for _ in range(world_size)]
        dist.all_gather(gi, image_features.detach().contiguous())
        dist.all_gather(gt, text_features.detach().contiguous())
        if not local_loss:      # keep the graph of this rank's own features
            gi[rank], gt[rank] = image_features, text_features
        all_image, all_text = torch.cat(gi, dim=0), torch.cat(gt, dim=0)
    return all_image, all_text


class ClipLoss(nn.Module):
    """loss.py:110-166."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1):
        super().__init__()
        self.local_loss, self.gather_with_grad, self.cache_labels = local_loss, gather_with_grad, cache_labels
        self.rank, self.world_size = rank, world_size
        self.labels: Dict[torch.device, torch.Tensor] = {}
        self.prev_num_logits = 0

    def forward(self, image_features, text_features, logit_scale):
        device = image_features.device
        if self.world_size > 1:
            all_image, all_text = gather_features(image_features, text_features, self.local_loss, self.gather_with_grad,
                                                  self.rank, self.world_size)
            if self.local_loss:
                logits_per_image = logit_scale * image_features @ all_text.T
                logits_per_text = logit_scale * text_features @ all_image.T
            else:
                logits_per_image = logit_scale * all_image @ all_text.T
                logits_per_text = logits_per_image.T
        else:
            logits_per_image = logit_scale * image_features @ text_features.T
            logits_per_text = logit_scale * text_features @ image_features.T
        n = logits_per_image.shape[0]
        if self.prev_num_logits != n or device not in self.labels:
            labels = torch.arange(n, device=device, dtype=torch.long)
            if self.world_size > 1 and self.local_loss:
                labels = labels + n * self.rank
            if self.cache_labels:
                self.labels[device], self.prev_num_logits = labels, n
        else:
            labels = self.labels[device]
        return (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_text, labels)) / 2


def average_gradients(grads, world: int) -> None:
    """One all-reduce over the concatenation of `grads`, averaged over the ranks, written back in place (what DDP's
    bucketed reducer does for the reference, TinyCLIP/src/training/main.py)."""
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    torch._foreach_copy_(grads, [piece.view_as(g) for piece, g in zip(flat.split([g.numel() for g in grads]), grads)])


class ClipTrainer:
    """One contrastive training step (TinyCLIP/src/training/train.py train_one_epoch, the plain CLIP
    branch): forward both towers, ClipLoss over the gathered features, backward, gradient average over
    ranks, AdamW with the reference's parameter groups and defaults (training/optimizer.py:22-52,
    params.py:8: lr 5e-4, betas (0.9, 0.98), eps 1e-6, wd 0.2), logit-scale clamp to [0, ln 100]
    (train.py:525-530)."""

    def __init__(self, model: CLIP, lr=5e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2, local_loss=True,
                 gather_with_grad=True, graph=False):
        """graph=True (single process): the whole step - both towers, loss, backward, AdamW, logit-scale clamp,
        ~6000 launches - is captured ONCE in a CUDA graph and replayed; the configuration is fixed, so unlike the
        supernet (a new subnet every step) nothing about the launch sequence changes from step to step."""
        self.model = model
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if self.world > 1:
            for p in model.parameters():
                dist.broadcast(p.data, src=0)
        self.use_graph = bool(graph) and self.world == 1
        named = list(model.named_parameters())
        skip = lambda n, p: p.ndim < 2 or "bn" in n or "ln" in n or "bias" in n or "logit_scale" in n
        self.opt = torch.optim.AdamW([
            dict(params=[p for n, p in named if skip(n, p)], weight_decay=0.0),
            dict(params=[p for n, p in named if not skip(n, p)], weight_decay=weight_decay)],
            lr=lr, betas=betas, eps=eps, fused=True, capturable=self.use_graph)
        self.loss = ClipLoss(local_loss=local_loss, gather_with_grad=gather_with_grad, cache_labels=True,
                             rank=self.rank, world_size=self.world)
        self.params = [p for _, p in named]
        self.stager = None
        self._graph = None

    def stage(self, images, texts):
        """Start the host -> device copy of the next batch on the side stream (staging.BatchStager)."""
        if self.stager is None:
            from .staging import BatchStager
            self.stager = BatchStager(self.params[0].device)
        return self.stager.stage(images, texts)

    def _step_eager(self, images, texts):
        self.opt.zero_grad(set_to_none=True)
        fi, ft, scale = self.model(images, texts)
        loss = self.loss(fi, ft, scale)
        loss.backward()
        if self.world > 1:
            average_gradients([p.grad for p in self.params if p.grad is not None], self.world)
        self.opt.step()
        ops.SHADOWS.invalidate()        # the fused optimiser writes parameters without a version bump
        with torch.no_grad():
            self.model.logit_scale.clamp_(0, math.log(100))
        return loss.detach()

    def _capture(self, images, texts):
        """Warm up (allocator pools, lazily set kernel attributes, optimizer state) WITHOUT advancing the training
        state, then record one step.  Returns False (and leaves the trainer eager) if the capture is refused."""
        self._g_images, self._g_texts = images.clone(), texts.clone()
        keep = [p.detach().clone() for p in self.params]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step_eager(self._g_images, self._g_texts)
            with torch.no_grad():
                for p, k in zip(self.params, keep):
                    p.copy_(k)
                for st in self.opt.state.values():      # moments and step counters back to a fresh optimizer
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
            ops.SHADOWS.invalidate()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.LAUNCHES[0]
        try:
            with torch.cuda.graph(graph):
                self._g_loss = self._step_eager(self._g_images, self._g_texts)
        except Exception as e:   # noqa: BLE001 - a refused capture must not take the training run down
            import warnings
            warnings.warn(f"cream_b200: CUDA-graph capture of the CLIP step failed ({e!r}); running eagerly")
            self.use_graph = False
            ops.SHADOWS.invalidate()
            return False
        self._graph_launches = _lib.LAUNCHES[0] - n0
        self._graph = graph
        return True

    def step(self, images, texts=None):
        """images: a batch tensor with `texts`, or the handle `stage` returned."""
        staged = images if texts is None else None
        if staged is not None:
            images, texts = staged.acquire()
        if self.use_graph and self._graph is not None and (images.shape != self._g_images.shape or texts.shape != self._g_texts.shape):
            self._graph = None          # another batch shape: the recorded launches no longer apply - capture again
        if self.use_graph and (self._graph is not None or self._capture(images, texts)):
            self._g_images.copy_(images, non_blocking=True)
            self._g_texts.copy_(texts, non_blocking=True)
            if staged is not None:
                staged.release()        # the slot has been copied into the graph's input buffers
            self._graph.replay()
            _lib.LAUNCHES[0] += self._graph_launches
            return self._g_loss.clone()
        loss = self._step_eager(images, texts)
        if staged is not None:
            staged.release()            # the text tower's backward reads the token ids once more
        return loss
