"""CHECKER INFRASTRUCTURE (never imported by the product path): loads the reference's OWN files —
unmodified — from `$CREAM_REFERENCE`, `/root/reference` (build container) or `baseline/_ref`
(staged by scripts/stage_reference.py; travels to the GPU box), with the few `sys.modules` shims
SURVEY.md 8c lists (torch._six, easydict, a timm-0.3.2-shaped namespace).

Two flavours of every load:
  * `over="reference"` — the reference resolves its own `model.module.*` / `rpe_ops`; this is the
    reference path itself (the GPU-side / CPU-side baseline and the parity target);
  * `over="cream"`     — `model.module.*`, `model.utils` and `rpe_ops` resolve to cream_b200's
    drop-ins while the reference's L2/L3 files (supernet_transformer.py, supernet_engine.py, irpe.py,
    rpe_vision_transformer.py) run unchanged on top: the drop-in claim of BASELINE.json's north_star.

Harness glue only: nothing here restates reference arithmetic.
"""
from __future__ import annotations

import collections.abc
import contextlib
import importlib
import importlib.util
import os
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent


def reference_root() -> Path | None:
    for cand in (os.environ.get("CREAM_REFERENCE"), "/root/reference", ROOT / "baseline" / "_ref"):
        if cand and (Path(cand) / "AutoFormer" / "model" / "supernet_transformer.py").exists():
            return Path(cand)
    return None


def available() -> bool:
    return reference_root() is not None


# ---------------------------------------------------------------------------------------------
# shims
# ---------------------------------------------------------------------------------------------
class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _DropPath(nn.Module):
    """timm 0.3.2 DropPath (per-sample stochastic depth)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.drop_prob or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = (keep + torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)).floor_()
        return x.div(keep) * mask


class _Mlp(nn.Module):
    """timm 0.3.2 vision_transformer.Mlp."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class _PatchEmbed(nn.Module):
    """timm 0.3.2 vision_transformer.PatchEmbed."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


def _accuracy(output, target, topk=(1,)):
    """timm.utils.accuracy."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.reshape(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0) * 100. / target.size(0) for k in topk]


def _unwrap_model(model):
    return model.module if hasattr(model, "module") else model


def install_shims() -> None:
    if "torch._six" not in sys.modules:
        six = types.ModuleType("torch._six")
        six.container_abcs = collections.abc
        sys.modules["torch._six"] = six
    if "easydict" not in sys.modules:
        ed = types.ModuleType("easydict")
        ed.EasyDict = _EasyDict
        sys.modules["easydict"] = ed
    if "timm" in sys.modules and not getattr(sys.modules["timm"], "_cream_shim", False):
        return  # a real timm is importable: use it

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m._cream_shim = True
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    mod("timm")
    mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225),
        Mixup=_Stub)
    mod("timm.utils", accuracy=_accuracy, ModelEma=_Stub)
    mod("timm.utils.model", unwrap_model=_unwrap_model)
    mod("timm.models")
    mod("timm.models.helpers", load_pretrained=lambda *a, **k: None, build_model_with_cfg=lambda *a, **k: None)
    mod("timm.models.layers", DropPath=_DropPath, to_2tuple=lambda x: x if isinstance(x, tuple) else (x, x),
        trunc_normal_=nn.init.trunc_normal_)
    mod("timm.models.resnet", resnet26d=None, resnet50d=None)
    mod("timm.models.registry", register_model=lambda f: f)
    mod("timm.models.vision_transformer", _cfg=lambda **k: k, default_cfgs={}, Mlp=_Mlp, PatchEmbed=_PatchEmbed,
        HybridEmbed=_Stub)


@contextlib.contextmanager
def _module_overrides(mapping: dict, drop_prefixes=()):
    """Temporarily install `mapping` into sys.modules (and hide modules under drop_prefixes); restore after."""
    saved = {}
    for k in list(sys.modules):
        if k in mapping or any(k == p or k.startswith(p + ".") for p in drop_prefixes):
            saved[k] = sys.modules.pop(k)
    sys.modules.update(mapping)
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        yield
    finally:
        sys.dont_write_bytecode = dont
        for k in list(sys.modules):
            if k in mapping or any(k == p or k.startswith(p + ".") for p in drop_prefixes):
                del sys.modules[k]
        sys.modules.update(saved)


def _exec_file(path: Path, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    try:
        spec.loader.exec_module(m)
    except BaseException:
        sys.modules.pop(name, None)
        raise
    return m


_CACHE: dict = {}


# ---------------------------------------------------------------------------------------------
# AutoFormer
# ---------------------------------------------------------------------------------------------
def autoformer(over: str = "reference"):
    """The reference's AutoFormer/model/supernet_transformer.py as a module object.
    over="reference": its own model.module.*; over="cream": cream_b200's drop-ins underneath."""
    key = ("autoformer", over)
    if key in _CACHE:
        return _CACHE[key]
    ref = reference_root()
    assert ref is not None, "reference not available (stage it with scripts/stage_reference.py)"
    install_shims()
    af = ref / "AutoFormer"
    if over == "cream":
        import cream_b200.autoformer.model as m
        import cream_b200.autoformer.model.module as mm
        import cream_b200.autoformer.model.utils as mu
        from cream_b200.autoformer.model.module import (Linear_super, embedding_super, layernorm_super,
                                                        multihead_super, qkv_super)
        mapping = {"model": m, "model.utils": mu, "model.module": mm}
        for sub in (Linear_super, embedding_super, layernorm_super, multihead_super, qkv_super):
            mapping["model.module." + sub.__name__.rsplit(".", 1)[1]] = sub
        with _module_overrides(mapping, drop_prefixes=("model",)):
            mod = _exec_file(af / "model" / "supernet_transformer.py", "cream_ref_supernet_transformer_over_cream")
    else:
        with _module_overrides({}, drop_prefixes=("model",)):
            sys.path.insert(0, str(af))
            try:
                importlib.invalidate_caches()
                mod = importlib.import_module("model.supernet_transformer")
            finally:
                sys.path.remove(str(af))
    _CACHE[key] = mod
    return mod


def supernet_engine():
    """The reference's AutoFormer/supernet_engine.py (train_one_epoch / evaluate / sample_configs)."""
    key = ("supernet_engine",)
    if key in _CACHE:
        return _CACHE[key]
    ref = reference_root()
    assert ref is not None
    install_shims()
    af = ref / "AutoFormer"
    with _module_overrides({}, drop_prefixes=("lib",)):
        sys.path.insert(0, str(af))
        try:
            importlib.invalidate_caches()
            mod = _exec_file(af / "supernet_engine.py", "cream_ref_supernet_engine")
        finally:
            sys.path.remove(str(af))
    _CACHE[key] = mod
    return mod


# ---------------------------------------------------------------------------------------------
# iRPE
# ---------------------------------------------------------------------------------------------
def reference_rpe_ops_cuda():
    """The reference's rpe_ops built WITH its CUDA kernels (oracle/build_ref.py) or None."""
    d = ROOT / "oracle" / "_ref" / "cuda"
    hits = sorted(d.glob("rpe_index_cpp*.so")) if d.exists() else []
    if not hits:
        return None
    if ("rpe_cuda",) not in _CACHE:
        spec = importlib.util.spec_from_file_location("rpe_index_cpp", hits[0])
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _CACHE[("rpe_cuda",)] = m
    return _CACHE[("rpe_cuda",)]


def irpe(over: str = "reference", flavour: str = "DeiT"):
    """The reference's irpe.py.  over: "reference" (pure-PyTorch gather fallback, irpe.py:8-15),
    "reference_cuda" (its own rpe_ops CUDA build), "cream" (cream_b200.rpe_ops drop-in).
    flavour: "DeiT" (iRPE/DeiT-with-iRPE/irpe.py) or "DETR" (models/rpe_attention/irpe.py)."""
    key = ("irpe", over, flavour)
    if key in _CACHE:
        return _CACHE[key]
    ref = reference_root()
    assert ref is not None
    install_shims()
    path = ref / "iRPE" / ("DeiT-with-iRPE/irpe.py" if flavour == "DeiT" else "DETR-with-iRPE/models/rpe_attention/irpe.py")
    mapping = {}
    if over == "cream":
        import cream_b200.rpe_ops as pkg
        import cream_b200.rpe_ops.rpe_index as ri
        import cream_b200.rpe_ops.rpe_index_cpp as cpp
        mapping = {"rpe_ops": pkg, "rpe_ops.rpe_index": ri, "rpe_index_cpp": cpp}
    elif over == "reference_cuda":
        cpp = reference_rpe_ops_cuda()
        assert cpp is not None, "oracle/_ref/cuda/rpe_index_cpp*.so not built"
        pkg = types.ModuleType("rpe_ops")
        pkg.__path__ = []
        with _module_overrides({"rpe_index_cpp": cpp}, drop_prefixes=("rpe_ops",)):
            ri = _exec_file(ref / "iRPE" / "DeiT-with-iRPE" / "rpe_ops" / "rpe_index.py", "cream_ref_rpe_index_py")
            sys.modules.pop("cream_ref_rpe_index_py", None)
        pkg.rpe_index = ri
        mapping = {"rpe_ops": pkg, "rpe_ops.rpe_index": ri, "rpe_index_cpp": cpp}
    else:
        blocker = types.ModuleType("rpe_ops")   # no attribute `rpe_index`, not a package: ImportError -> fallback
        mapping = {"rpe_ops": blocker}
    import warnings
    with _module_overrides(mapping, drop_prefixes=("rpe_ops", "rpe_index_cpp")), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = _exec_file(path, f"cream_ref_irpe_{over}_{flavour}")
    want = over != "reference"
    assert (mod.RPEIndexFunction is not None) == want, f"irpe.py picked the wrong rpe_ops for over={over}"
    _CACHE[key] = mod
    return mod


def rpe_vision_transformer(over: str = "reference"):
    """The reference's rpe_vision_transformer.py (RPEAttention / RPEBlock / VisionTransformer) bound to
    the irpe module of the requested flavour."""
    key = ("rpe_vit", over)
    if key in _CACHE:
        return _CACHE[key]
    ref = reference_root()
    mod_irpe = irpe(over)
    with _module_overrides({"irpe": mod_irpe}):
        mod = _exec_file(ref / "iRPE" / "DeiT-with-iRPE" / "rpe_vision_transformer.py", f"cream_ref_rpe_vit_{over}")
    mod.irpe = mod_irpe
    _CACHE[key] = mod
    return mod


# ---------------------------------------------------------------------------------------------
# TinyCLIP
# ---------------------------------------------------------------------------------------------
def open_clip_model():
    """The reference's TinyCLIP/src/open_clip/model.py as a module, without running the package
    __init__ (which pulls ftfy / regex through the tokenizer): its sibling imports (.timm_model, .utils,
    .resnet, .l0module) resolve to name-only stubs - the attention path needs none of them."""
    key = ("open_clip_model",)
    if key in _CACHE:
        return _CACHE[key]
    ref = reference_root()
    assert ref is not None
    path = ref / "TinyCLIP" / "src" / "open_clip" / "model.py"
    pkg = types.ModuleType("cream_ref_open_clip")
    pkg.__path__ = []

    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    def sub(name, **attrs):
        m = types.ModuleType("cream_ref_open_clip." + name)
        for k, v in attrs.items():
            setattr(m, k, v)
        return m
    mapping = {"cream_ref_open_clip": pkg,
               "cream_ref_open_clip.timm_model": sub("timm_model", TimmModel=_Stub),
               "cream_ref_open_clip.utils": sub("utils", freeze_batch_norm_2d=lambda *a, **k: None,
                                                to_2tuple=lambda x: x if isinstance(x, tuple) else (x, x)),
               "cream_ref_open_clip.resnet": sub("resnet", ModifiedResNet=_Stub),
               "cream_ref_open_clip.l0module": sub("l0module", L0Module=_Stub)}
    with _module_overrides(mapping):
        spec = importlib.util.spec_from_file_location("cream_ref_open_clip.model", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["cream_ref_open_clip.model"] = mod
        try:
            spec.loader.exec_module(mod)
        finally:
            sys.modules.pop("cream_ref_open_clip.model", None)
    _CACHE[key] = mod
    return mod


def open_clip_loss():
    """The reference's TinyCLIP/src/open_clip/loss.py as a module (it imports only torch)."""
    key = ("open_clip_loss",)
    if key in _CACHE:
        return _CACHE[key]
    ref = reference_root()
    assert ref is not None
    path = ref / "TinyCLIP" / "src" / "open_clip" / "loss.py"
    spec = importlib.util.spec_from_file_location("cream_ref_open_clip_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _CACHE[key] = mod
    return mod
