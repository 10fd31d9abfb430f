"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports /root/reference (read-only) with two sys.modules shims (torch._six, easydict) and
a stub `timm` namespace so that `RPEAttention` can be imported as-is; writes small .npz
files next to this script.  /root/reference does not exist on the GPU box, so tests only
ever read the committed fixtures.  Inputs/weights are NOT stored: they are regenerated
from numpy PCG64 seeds by oracle.vit_oracle.init_params / the helpers below.
"""
from __future__ import annotations

import collections.abc
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path(os.environ.get("CREAM_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True

from oracle import vit_oracle as vo  # noqa: E402


# ---------------------------------------------------------------------------------------
# shims (harness glue, not reference code)
# ---------------------------------------------------------------------------------------
def install_shims():
    six = types.ModuleType("torch._six")
    six.container_abcs = collections.abc
    sys.modules["torch._six"] = six

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in {**(d or {}), **kw}.items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __setitem__ = __setattr__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

    ed = types.ModuleType("easydict")
    ed.EasyDict = EasyDict
    sys.modules["easydict"] = ed

    # stub timm: only names, enough to import rpe_vision_transformer.py and use RPEAttention
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Stub(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    ident = lambda *a, **k: (lambda f: f) if not a or not callable(a[0]) else a[0]
    mod("timm")
    mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    mod("timm.models")
    mod("timm.models.helpers", load_pretrained=lambda *a, **k: None, build_model_with_cfg=lambda *a, **k: None)
    mod("timm.models.layers", DropPath=_Stub, to_2tuple=lambda x: (x, x),
        trunc_normal_=torch.nn.init.trunc_normal_)
    mod("timm.models.resnet", resnet26d=None, resnet50d=None)
    mod("timm.models.registry", register_model=lambda f: f)
    mod("timm.models.vision_transformer", _cfg=lambda **k: k, default_cfgs={}, Mlp=_Stub, PatchEmbed=_Stub,
        HybridEmbed=_Stub)


def rand(shape, seed, scale=1.0):
    return torch.from_numpy((np.random.default_rng(seed).standard_normal(shape, dtype=np.float32) * scale))


def summarize(t: torch.Tensor, full_limit=20000):
    """Full tensor when small, else (sum, abs-sum, strided sample)."""
    a = t.detach().double().numpy()
    if a.size <= full_limit:
        return {"full": a.astype(np.float32)}
    flat = a.reshape(-1)
    return {"sum": np.array(flat.sum()), "abssum": np.array(np.abs(flat).sum()),
            "sample": flat[::101].astype(np.float32)}


# ---------------------------------------------------------------------------------------
def golden_index_tables():
    sys.path.insert(0, str(REF / "iRPE" / "DeiT-with-iRPE"))
    import irpe  # reference, unmodified
    sys.path.insert(0, str(REF / "AutoFormer"))
    from model.module.multihead_super import RelativePosition2D_super

    out = {}
    # piecewise_index over a wide integer range and the float path
    xs = torch.arange(-700, 701)
    for ratio in (1.9, 1.0, 2.5, 3.3):
        a, b, g = 1 * ratio, 2 * ratio, 8 * ratio
        out[f"pw_int_{ratio}"] = irpe.piecewise_index(xs, a, b, g, torch.long).numpy()
        xf = torch.arange(0, 60, dtype=torch.float32)
        out[f"pw_float_{ratio}"] = irpe.piecewise_index(xf, a, b, g, torch.long).numpy()
    out["pw_x"] = xs.numpy()
    # bucket ids
    methods = {"euc": irpe.METHOD.EUCLIDEAN, "quant": irpe.METHOD.QUANT, "product": irpe.METHOD.PRODUCT,
               "rows": irpe.METHOD.CROSS_ROWS, "cols": irpe.METHOD.CROSS_COLS}
    for mname, mid in methods.items():
        for (h, w, skip, ratio) in [(14, 14, 1, 1.9), (7, 7, 0, 1.9), (5, 9, 2, 1.9), (14, 14, 1, 3.3), (24, 24, 1, 1.9)]:
            irpe.BUCKET_IDS_BUF.clear()
            ids, nb = irpe.get_bucket_ids_2d(method=mid, height=h, width=w, skip=skip, alpha=1 * ratio,
                                             beta=2 * ratio, gamma=8 * ratio, dtype=torch.long)
            out[f"ids_{mname}_{h}_{w}_{skip}_{ratio}"] = ids.numpy().astype(np.int32)
            out[f"nb_{mname}_{h}_{w}_{skip}_{ratio}"] = np.array(nb)
    # AutoFormer 2D relative position indices: recover them through one-hot tables
    for grid in (14, 4, 7):
        n = grid * grid + 1
        m = RelativePosition2D_super(30, 14)
        with torch.no_grad():
            m.embeddings_table_v.copy_(torch.eye(30))
            m.embeddings_table_h.zero_()
        m.set_sample_config(30)
        iv = m(n, n).argmax(-1)
        with torch.no_grad():
            m.embeddings_table_h.copy_(torch.eye(30))
            m.embeddings_table_v.zero_()
        m.set_sample_config(30)
        ih = m(n, n).argmax(-1)
        out[f"af_idx_v_{grid}"] = iv.numpy().astype(np.int32)
        out[f"af_idx_h_{grid}"] = ih.numpy().astype(np.int32)
    np.savez_compressed(HERE / "index_tables.npz", **out)
    print("index_tables.npz", len(out), "arrays")


MICRO_SPECS = {
    # name: (spec, batch, configs)
    "micro17": (vo.SupernetSpec(embed_dim=128, depth=3, num_heads=2, mlp_ratio=4.0, img_size=64), 3, [
        {"layer_num": 2, "embed_dim": [64] * 2, "num_heads": [1, 1], "mlp_ratio": [3.5, 4.0]},
        {"layer_num": 3, "embed_dim": [128] * 3, "num_heads": [2, 1, 2], "mlp_ratio": [4.0, 3.0, 3.5]},
        {"layer_num": 3, "embed_dim": [96] * 3, "num_heads": [2, 2, 1], "mlp_ratio": [3.0, 3.5, 4.0]},
    ]),
    "micro197": (vo.SupernetSpec(embed_dim=192, depth=3, num_heads=3, mlp_ratio=4.0, img_size=224), 2, [
        {"layer_num": 2, "embed_dim": [128] * 2, "num_heads": [2, 3], "mlp_ratio": [3.5, 4.0]},
        {"layer_num": 3, "embed_dim": [192] * 3, "num_heads": [3, 3, 2], "mlp_ratio": [4.0, 3.5, 3.0]},
    ]),
}


def golden_supernet():
    sys.path.insert(0, str(REF / "AutoFormer"))
    from model.supernet_transformer import Vision_TransformerSuper  # reference, unmodified

    out = {}
    for name, (spec, batch, configs) in MICRO_SPECS.items():
        net = Vision_TransformerSuper(img_size=spec.img_size, patch_size=spec.patch_size, embed_dim=spec.embed_dim,
                                      depth=spec.depth, num_heads=spec.num_heads, mlp_ratio=spec.mlp_ratio,
                                      qkv_bias=True, drop_rate=0.0, drop_path_rate=0.0, gp=True,
                                      num_classes=spec.num_classes, max_relative_position=14,
                                      relative_position=True, change_qkv=True, abs_pos=True)
        sd = vo.init_params(spec, seed=7)
        ref_shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert ref_shapes == {k: tuple(v.shape) for k, v in sd.items()}, "oracle param_shapes != reference state_dict"
        net.load_state_dict(sd)
        net.train()
        images = rand((batch, 3, spec.img_size, spec.img_size), seed=11)
        targets = torch.from_numpy(np.random.default_rng(13).integers(0, spec.num_classes, batch))
        for ci, cfg in enumerate(configs):
            net.zero_grad(set_to_none=True)
            net.set_sample_config(cfg)
            logits = net(images)
            loss = torch.nn.functional.cross_entropy(logits, targets)
            loss.backward()
            key = f"{name}_c{ci}"
            out[f"{key}_logits"] = logits.detach().numpy()
            out[f"{key}_loss"] = np.array(loss.item())
            out[f"{key}_numel"] = np.array(net.get_sampled_params_numel(cfg))
            none_names = []
            for pn, p in net.named_parameters():
                if p.grad is None:
                    none_names.append(pn)
                    continue
                for sk, sv in summarize(p.grad).items():
                    out[f"{key}_grad_{pn}_{sk}"] = sv
            out[f"{key}_none"] = np.array(none_names)
    np.savez_compressed(HERE / "supernet_micro.npz", **out)
    print("supernet_micro.npz", len(out), "arrays")


IRPE_CASES = {
    # name: (rpe_on, mode, shared_head, method, C, heads, grid)
    "k_ctx_shared": ("k", "ctx", True, "product", 128, 2, 14),     # BASELINE config 2 kind
    "qkv_ctx_perhead": ("qkv", "ctx", False, "product", 128, 2, 7),
    "qk_bias": ("qk", "bias", False, "euc", 64, 2, 7),
    "k_ctx_quant": ("k", "ctx", True, "quant", 64, 1, 5),
    "kv_ctx_cross": ("kv", "ctx", True, "cross", 128, 2, 7),       # iRPE_Cross: rows + cols tables
}


def golden_irpe_attention():
    sys.path.insert(0, str(REF / "iRPE" / "DeiT-with-iRPE"))
    import irpe
    from rpe_vision_transformer import RPEAttention  # reference, unmodified (timm stubbed)

    out = {}
    for name, (rpe_on, mode, shared, method, C, heads, grid) in IRPE_CASES.items():
        irpe.BUCKET_IDS_BUF.clear()
        cfg = irpe.get_rpe_config(ratio=1.9, method=method, mode=mode, shared_head=shared, skip=1, rpe_on=rpe_on)
        attn = RPEAttention(C, num_heads=heads, qkv_bias=True, rpe_config=cfg)
        N = grid * grid + 1
        B = 2
        seed = 100
        with torch.no_grad():
            for pn, p in attn.named_parameters():
                seed += 1
                p.copy_(rand(tuple(p.shape), seed, 0.3 if "lookup" in pn else 0.08))
        x = rand((B, N, C), 99).requires_grad_(True)
        gy = rand((B, N, C), 98)
        y = attn(x)
        y.backward(gy)
        for sk, sv in summarize(y).items():
            out[f"{name}_y_{sk}"] = sv
        for sk, sv in summarize(x.grad).items():
            out[f"{name}_gx_{sk}"] = sv
        for pn, p in attn.named_parameters():
            out[f"{name}_shape_{pn}"] = np.array(p.shape)
            for sk, sv in summarize(p.grad).items():
                out[f"{name}_grad_{pn}_{sk}"] = sv
    np.savez_compressed(HERE / "irpe_attention.npz", **out)
    print("irpe_attention.npz", len(out), "arrays")


TINYVIT_CASES = {
    # name: (dim, key_dim, heads, attn_ratio, resolution, batch)
    "win7": (64, 32, 2, 1, (7, 7), 6),        # 7 x 7 windows (TinyViT stages 2-3; batch = B * windows)
    "full14": (128, 32, 4, 1, (14, 14), 2),   # 14 x 14, the last-stage resolution at 224 / 16
}


def golden_tinyvit_attention():
    """TinyViT/models/tiny_vit.py:215-286 `Attention` (LayerNorm -> qkv -> per-head bias gather ->
    softmax -> proj), imported unmodified (timm stubbed)."""
    sys.path.insert(0, str(REF / "TinyViT"))
    from models.tiny_vit import Attention  # reference, unmodified

    out = {}
    for name, (dim, key_dim, heads, ratio, res, B) in TINYVIT_CASES.items():
        attn = Attention(dim, key_dim, heads, attn_ratio=ratio, resolution=res)
        attn.train()
        seed = 300
        with torch.no_grad():
            for pn, p in attn.named_parameters():
                seed += 1
                if pn == "norm.weight":
                    p.copy_(1.0 + rand(tuple(p.shape), seed, 0.1))
                else:
                    p.copy_(rand(tuple(p.shape), seed, 0.5 if "attention_biases" in pn else 0.08))
        N = res[0] * res[1]
        x = rand((B, N, dim), 299).requires_grad_(True)
        gy = rand((B, N, dim), 298)
        y = attn(x)
        y.backward(gy)
        out[f"{name}_idxs"] = attn.attention_bias_idxs.numpy()
        for sk, sv in summarize(y).items():
            out[f"{name}_y_{sk}"] = sv
        for sk, sv in summarize(x.grad).items():
            out[f"{name}_gx_{sk}"] = sv
        for pn, p in attn.named_parameters():
            out[f"{name}_shape_{pn}"] = np.array(p.shape)
            for sk, sv in summarize(p.grad).items():
                out[f"{name}_grad_{pn}_{sk}"] = sv
    np.savez_compressed(HERE / "tinyvit_attention.npz", **out)
    print("tinyvit_attention.npz", len(out), "arrays")


def golden_rpe_index():
    """The reference's own self-check shape (rpe_ops/rpe_index.py:59-100), through the
    reference C++ op compiled by oracle/build_ref.py."""
    sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
    try:
        import rpe_index_cpp  # noqa
    except ImportError:
        print("oracle/_ref/rpe_index_cpp not built; run python oracle/build_ref.py first")
        return
    assert rpe_index_cpp.version() == "1.2.0"
    B, H, L, nb = 4, 3, 50, 50
    x = rand((B, H, L, nb), 5)
    idx = torch.from_numpy(np.random.default_rng(6).integers(0, nb, (L, L)).astype(np.int32))
    y = rpe_index_cpp.forward_cpu(x, idx)
    gy = rand((B, H, L, L), 8)
    gx = torch.zeros_like(x)
    rpe_index_cpp.backward_cpu(gx, gy, idx)
    np.savez_compressed(HERE / "rpe_index.npz", y=y.numpy(), gx=gx.numpy())
    print("rpe_index.npz")


if __name__ == "__main__":
    assert REF.exists(), f"{REF} not found: golden fixtures can only be regenerated in the build container"
    torch.manual_seed(0)
    install_shims()
    golden_index_tables()
    golden_supernet()
    golden_irpe_attention()
    golden_tinyvit_attention()
    golden_rpe_index()
