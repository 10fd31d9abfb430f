"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol declared in
include/cream_b200.h, the host-side integer tables are bit exact against the reference's
fixtures, and the host-side module mirror has the reference's surface (names, shapes,
sampled-parameter counts, error behaviour)."""
import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import rel_index, vit_oracle as vo

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))
from make_golden import MICRO_SPECS  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from cream_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from cream_b200 import _lib
    header = (ROOT / "include" / "cream_b200.h").read_text()
    declared = set(re.findall(r"\b(cream_[a-z0-9_]+)\s*\(", header))
    declared -= {"cream_gemm_desc", "cream_attn_desc"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/cream_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), f"binding table out of sync: {declared ^ set(_lib.SIGNATURES)}"
    assert lib.cream_rpe_index_version() == b"1.2.0"          # rpe_index.py:5-8 asserts this
    assert b"sm_100a" in lib.cream_version()


def test_struct_layouts_match_header():
    from cream_b200 import _lib
    # natural alignment, same field order as the header: spot-check offsets that matter
    assert _lib.GemmDesc.a.offset == 16 and _lib.GemmDesc.b.offset == 40
    assert ctypes.sizeof(_lib.GemmDesc) % 8 == 0 and ctypes.sizeof(_lib.AttnDesc) % 8 == 0
    assert _lib.AttnDesc.qkv.offset == 24


def test_host_autoformer_index_bit_exact(lib, golden_dir):
    t = np.load(golden_dir / "index_tables.npz")
    for grid in (14, 4, 7):
        n = grid * grid + 1
        iv = np.empty((n, n), np.int32)
        ih = np.empty((n, n), np.int32)
        assert lib.cream_autoformer_rel_index_host(grid, 14, iv.ctypes.data, ih.ctypes.data) == 0
        np.testing.assert_array_equal(iv, t[f"af_idx_v_{grid}"])
        np.testing.assert_array_equal(ih, t[f"af_idx_h_{grid}"])


@pytest.mark.parametrize("mname,mid", [("euc", 0), ("quant", 1), ("product", 3), ("rows", 41), ("cols", 42)])
def test_host_irpe_bucket_ids_bit_exact(lib, golden_dir, mname, mid):
    t = np.load(golden_dir / "index_tables.npz")
    for (h, w, skip, ratio) in [(14, 14, 1, 1.9), (7, 7, 0, 1.9), (5, 9, 2, 1.9), (14, 14, 1, 3.3), (24, 24, 1, 1.9)]:
        n = skip + h * w
        out = np.empty((n, n), np.int32)
        nb = ctypes.c_int(0)
        assert lib.cream_irpe_bucket_ids_host(mid, h, w, skip, 1 * ratio, 2 * ratio, 8 * ratio, out.ctypes.data,
                                              ctypes.byref(nb)) == 0
        key = f"{mname}_{h}_{w}_{skip}_{ratio}"
        assert nb.value == int(t["nb_" + key])
        np.testing.assert_array_equal(out, t["ids_" + key])
        ids, nb2 = rel_index.irpe_bucket_ids(mid, h, w, skip, 1 * ratio, 2 * ratio, 8 * ratio)
        np.testing.assert_array_equal(out, ids)


def test_host_table_argument_errors(lib):
    assert lib.cream_autoformer_rel_index_host(0, 14, None, None) != 0
    nb = ctypes.c_int(0)
    out = np.empty((4, 4), np.int32)
    assert lib.cream_irpe_bucket_ids_host(7, 2, 2, 0, 1.9, 3.8, 15.2, out.ctypes.data, ctypes.byref(nb)) != 0


@pytest.mark.parametrize("name", list(MICRO_SPECS))
def test_module_mirror_surface(golden_dir, name):
    from cream_b200.autoformer.model.supernet_transformer import Vision_TransformerSuper
    g = np.load(golden_dir / "supernet_micro.npz")
    spec, batch, configs = MICRO_SPECS[name]
    net = Vision_TransformerSuper(img_size=spec.img_size, embed_dim=spec.embed_dim, depth=spec.depth,
                                  num_heads=spec.num_heads, mlp_ratio=spec.mlp_ratio, qkv_bias=True, gp=True,
                                  relative_position=True, change_qkv=True, max_relative_position=14)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert shapes == vo.param_shapes(spec), "state_dict names/shapes must equal the reference's"
    net.load_state_dict(vo.init_params(spec, seed=7))
    assert net.no_weight_decay() == {'pos_embed', 'cls_token', 'rel_pos_embed'}
    assert isinstance(net.head, torch.nn.Linear) and isinstance(net.norm, torch.nn.LayerNorm)
    for ci, cfg in enumerate(configs):
        assert net.get_sampled_params_numel(cfg) == int(g[f"{name}_c{ci}_numel"])
        assert net.get_sampled_params_numel(cfg) == vo.sampled_param_count(cfg, spec)
        assert net.get_complexity(net.patch_embed_super.num_patches) > 0
        assert net.blocks[cfg["layer_num"] - 1].is_identity_layer is False
        if cfg["layer_num"] < spec.depth:
            assert net.blocks[-1].is_identity_layer is True
    with pytest.raises(RuntimeError):
        net(torch.randn(1, 3, spec.img_size, spec.img_size))      # CPU tensors: fail loudly, no fallback


def test_sampled_param_names_exclude_identity_layers():
    from cream_b200 import engine
    geo = engine.SupernetGeometry(embed_dim=128, depth=3, num_heads=2, mlp_ratio=4.0, img_size=64)
    names = engine.sampled_param_names(geo, {"layer_num": 2, "embed_dim": [64, 64], "num_heads": [1, 1],
                                             "mlp_ratio": [3.5, 4.0]})
    assert not any(n.startswith("blocks.2.") for n in names)
    assert set(names) <= set(vo.param_shapes(vo.SupernetSpec(128, 3, 2, 4.0, img_size=64)))
    with pytest.raises(AssertionError):
        engine.validate_config(geo, {"layer_num": 4, "embed_dim": [64] * 4, "num_heads": [1] * 4, "mlp_ratio": [4.0] * 4})


def test_rpe_ops_drop_in_surface():
    import cream_b200.rpe_ops.rpe_index as ri
    import rpe_index_cpp  # registered under the reference's top-level module name
    assert rpe_index_cpp.version() == "1.2.0" and ri.EXPECTED_VERSION == "1.2.0"
    for fn in ("forward_cpu", "backward_cpu", "forward_gpu", "backward_gpu", "version"):
        assert hasattr(rpe_index_cpp, fn)
    with pytest.raises(RuntimeError):
        ri.RPEIndexFunction.apply(torch.randn(1, 1, 3, 4), torch.zeros(3, 3, dtype=torch.int32))


REF = Path("/root/reference")


@pytest.mark.skipif(not REF.exists(), reason="reference checkout only exists in the build container")
def test_reference_model_file_runs_unchanged_on_the_drop_in_modules():
    """Import the reference's OWN supernet_transformer.py with `model.module.*` / `model.utils`
    resolved to cream_b200's drop-ins: construction, state_dict, set_sample_config and the
    sampled-parameter count must all work untouched (forward needs the GPU)."""
    import importlib.util
    import types
    import cream_b200.autoformer.model as m
    import cream_b200.autoformer.model.module as mm
    import cream_b200.autoformer.model.utils as mu
    from cream_b200.autoformer.model.module import Linear_super, embedding_super, layernorm_super, multihead_super, qkv_super
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "model" or k.startswith("model.")}
    try:
        sys.modules["model"] = m
        sys.modules["model.utils"] = mu
        sys.modules["model.module"] = mm
        for sub in (Linear_super, embedding_super, layernorm_super, multihead_super, qkv_super):
            sys.modules["model.module." + sub.__name__.rsplit(".", 1)[1]] = sub
        spec_ = importlib.util.spec_from_file_location("ref_supernet_transformer",
                                                       REF / "AutoFormer" / "model" / "supernet_transformer.py")
        ref_mod = importlib.util.module_from_spec(spec_)
        sys.dont_write_bytecode = True
        spec_.loader.exec_module(ref_mod)
        spec, batch, configs = MICRO_SPECS["micro17"]
        net = ref_mod.Vision_TransformerSuper(img_size=64, embed_dim=128, depth=3, num_heads=2, mlp_ratio=4.0,
                                              qkv_bias=True, gp=True, relative_position=True, change_qkv=True,
                                              max_relative_position=14)
        assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == vo.param_shapes(spec)
        for cfg in configs:
            assert net.get_sampled_params_numel(cfg) == vo.sampled_param_count(cfg, spec)
    finally:
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


def test_adapter_mirrors_surface_and_fail_loudly(golden_dir):
    """TinyViT / TinyCLIP / iRPE attention mirrors: the reference's parameter names and shapes, and
    no CPU path."""
    from cream_b200.clip_attention import ClipAttention
    from cream_b200.irpe_attention import RPEAttention
    from cream_b200.tinyvit_attention import Attention
    from make_golden import IRPE_CASES, TINYVIT_CASES
    g = np.load(golden_dir / "tinyvit_attention.npz")
    for name, (dim, key_dim, heads, ratio, res, B) in TINYVIT_CASES.items():
        m = Attention(dim, key_dim, heads, attn_ratio=ratio, resolution=res)
        want = {k[len(name) + 7:]: tuple(int(v) for v in g[k]) for k in g.files if k.startswith(name + "_shape_")}
        assert {k: tuple(v.shape) for k, v in m.named_parameters()} == want
        np.testing.assert_array_equal(m.attention_bias_idxs.numpy(), g[f"{name}_idxs"])
        with pytest.raises(RuntimeError):
            m(torch.randn(1, res[0] * res[1], dim))
    mha = torch.nn.MultiheadAttention(128, 2)
    clip = ClipAttention(128, 2)
    assert {k: tuple(v.shape) for k, v in clip.named_parameters()} == {k: tuple(v.shape) for k, v in mha.named_parameters()}
    with pytest.raises(RuntimeError):
        clip(torch.randn(5, 2, 128))
    gi = np.load(golden_dir / "irpe_attention.npz")
    for name, (rpe_on, mode, shared, method, C, heads, grid) in IRPE_CASES.items():
        if C // heads != 64:
            continue
        m = RPEAttention(C, num_heads=heads, qkv_bias=True, rpe_on=rpe_on, method=method,
                         mode="bias" if mode == "bias" else "contextual", shared_head=shared)
        want = {k[len(name) + 7:]: tuple(int(v) for v in gi[k]) for k in gi.files if k.startswith(name + "_shape_")}
        assert {k: tuple(v.shape) for k, v in m.named_parameters()} == want, name


def test_host_irpe_bucket_ids_property(lib):
    """Property test of the bit-exact integer contract: the library's host tables equal the numpy
    restatement of irpe.py for random grids, skips and ratios (incl. non-square grids, the
    height/width arguments DETR passes, rpe_attention_function.py:327-376)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(mid=st.sampled_from([0, 1, 3, 41, 42]), h=st.integers(1, 16), w=st.integers(1, 16),
           skip=st.integers(0, 2), ratio=st.floats(1.0, 4.0, allow_nan=False, width=32))
    def check(mid, h, w, skip, ratio):
        n = skip + h * w
        out = np.empty((n, n), np.int32)
        nb = ctypes.c_int(0)
        rc = lib.cream_irpe_bucket_ids_host(mid, h, w, skip, 1 * ratio, 2 * ratio, 8 * ratio, out.ctypes.data,
                                            ctypes.byref(nb))
        assert rc == 0
        ids, nb2 = rel_index.irpe_bucket_ids(mid, h, w, skip, 1 * ratio, 2 * ratio, 8 * ratio)
        assert nb.value == nb2
        np.testing.assert_array_equal(out, ids)

    check()


def test_host_autoformer_rel_index_property(lib):
    """Library host table vs the numpy restatement for every (grid, max_rel) in a range that includes
    binding clamps (max_rel < grid - 1) — bit exact."""
    for grid in range(1, 17):
        for max_rel in (1, 2, 5, grid - 1 if grid > 1 else 1, 14, 20):
            n = grid * grid + 1
            iv, ih = np.empty((n, n), np.int32), np.empty((n, n), np.int32)
            assert lib.cream_autoformer_rel_index_host(grid, max_rel, iv.ctypes.data, ih.ctypes.data) == 0
            ov, oh = rel_index.autoformer_rel_index(grid, max_rel)
            np.testing.assert_array_equal(iv, ov)
            np.testing.assert_array_equal(ih, oh)


@pytest.mark.skipif(not REF.exists(), reason="reference checkout only exists in the build container")
def test_non_square_bucket_ids_match_the_detr_copy_of_irpe():
    """height / width arguments (iRPE/DETR-with-iRPE/models/rpe_attention/irpe.py, used by
    rpe_attention_function.py:327-376): the numpy restatement and the library's host table vs the
    reference function imported in place."""
    import ctypes
    from oracle import refload
    from cream_b200 import _lib
    lib = _lib.load()
    irpe = refload.irpe("reference", "DETR")
    methods = {"product": (irpe.METHOD.PRODUCT, rel_index.PRODUCT), "euc": (irpe.METHOD.EUCLIDEAN, rel_index.EUCLIDEAN),
               "quant": (irpe.METHOD.QUANT, rel_index.QUANT), "rows": (irpe.METHOD.CROSS_ROWS, rel_index.CROSS_ROWS),
               "cols": (irpe.METHOD.CROSS_COLS, rel_index.CROSS_COLS)}
    for name, (ref_id, my_id) in methods.items():
        for (h, w, skip) in ((8, 12, 1), (11, 13, 2), (6, 16, 0), (3, 25, 1)):
            irpe.BUCKET_IDS_BUF.clear()
            ids, nb = irpe.get_bucket_ids_2d(method=ref_id, height=h, width=w, skip=skip, alpha=1.9, beta=3.8, gamma=15.2,
                                             dtype=torch.long)
            mine, nb2 = rel_index.irpe_bucket_ids(my_id, h, w, skip, 1.9, 3.8, 15.2)
            assert nb == nb2, (name, h, w, skip)
            np.testing.assert_array_equal(ids.numpy(), mine)
            n = skip + h * w
            out = np.empty((n, n), np.int32)
            nbc = ctypes.c_int(0)
            assert lib.cream_irpe_bucket_ids_host(my_id, h, w, skip, 1.9, 3.8, 15.2, out.ctypes.data, ctypes.byref(nbc)) == 0
            assert nbc.value == nb
            np.testing.assert_array_equal(out, mine)


def test_short_sequence_packing_host_logic(golden_dir):
    """Host side of cream_attn_desc.block_len: which (batch, tokens, block) the TinyCLIP towers hand to the attention kernel
    (clip._attn_packing) and the offset tables TinyViT's windows use for the in-kernel bias gather (one window / two windows laid
    end to end; more than 64 distinct offsets fall back to the dense logit term)."""
    from cream_b200 import clip, ops
    from cream_b200.tinyvit_attention import Attention
    assert clip._attn_packing(128, 50, False) == (64, 100, 50)        # image tower: two items per 128-row tile
    assert clip._attn_packing(127, 50, False) == (127, 50, 0)         # odd batch: one item per tile
    assert clip._attn_packing(128, 77, True) == (128, 77, 0)          # text tower: 154 tokens do not fit a tile
    assert clip._attn_packing(4, 64, True) == (2, 128, 64)
    win = Attention(64, 32, 2, attn_ratio=1, resolution=(7, 7))
    ids = win.attention_bias_idxs.numpy()
    assert win._ids1 is not None and int(win._ids1.max()) + 1 == 49 <= ops.NB_PACK
    np.testing.assert_array_equal(win._ids1, ids)
    assert win._ids2.shape == (98, 98)
    for a in range(2):
        for b in range(2):
            np.testing.assert_array_equal(win._ids2[49 * a:49 * (a + 1), 49 * b:49 * (b + 1)], ids)
    full = Attention(128, 32, 4, attn_ratio=1, resolution=(14, 14))
    assert full._ids1 is None and full._ids2 is None                   # 196 offsets: dense logit term


def test_shadow_cache_bookkeeping(monkeypatch):
    """ShadowCache host logic with the cast kernels stubbed out (no GPU): one cast per parameter version, a hit for another
    wrapper of the same parameter (what autograd hands to backward), re-cast after invalidate(), the entry pins the parameter's
    storage while the parameter lives, and the purge drops it once the parameter object is gone."""
    import gc
    from cream_b200 import _lib, ops
    calls = []

    class _Stub:
        def cream_shadow_cast(self, *a):
            calls.append("cast")
            return 0

        def cream_shadow_qkv(self, *a):
            calls.append("qkv")
            return 0

    monkeypatch.setattr(_lib, "load", lambda: _Stub())
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    cache = ops.ShadowCache()
    w = torch.nn.Parameter(torch.randn(12, 20))
    s1 = cache.get(w)
    assert calls == ["cast"] and tuple(s1.shape) == (12, 24) and s1.dtype == torch.bfloat16
    assert cache.get(w) is s1 and cache.get(w.detach().requires_grad_()) is s1 and calls == ["cast"]   # same storage, same version
    with torch.no_grad():
        w.add_(1.0)                                           # version bump: re-cast into the SAME buffer
    assert cache.get(w) is s1 and calls == ["cast", "cast"]
    cache.invalidate(w)
    assert cache.get(w) is s1 and len(calls) == 3
    sq = cache.get(w, qkv=True)
    assert calls[-1] == "qkv" and sq is not s1
    (key, ent), = [(k, e) for k, e in cache._store.items() if not k[1]]
    assert ent[2]._cdata == w.untyped_storage()._cdata and ent[3]() is w
    cache._purge()
    assert len(cache._store) == 2                              # parameter alive: nothing dropped
    del w, ent
    gc.collect()
    cache._purge()
    assert len(cache._store) == 0                              # parameter gone: storage and shadow released
