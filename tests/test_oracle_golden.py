"""The oracle (oracle/) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only; this is what pins the oracle."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import rel_index, vit_oracle as vo
from tests.helpers import check_summary, rand

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def tables(golden_dir):
    return np.load(golden_dir / "index_tables.npz")


@pytest.mark.parametrize("ratio", [1.9, 1.0, 2.5, 3.3])
def test_piecewise_index_bit_exact(tables, ratio):
    xs = tables["pw_x"]
    got = rel_index.piecewise_index(xs, 1 * ratio, 2 * ratio, 8 * ratio)
    np.testing.assert_array_equal(got, tables[f"pw_int_{ratio}"])
    xf = np.arange(0, 60, dtype=np.float32)
    np.testing.assert_array_equal(rel_index.piecewise_index(xf, 1 * ratio, 2 * ratio, 8 * ratio),
                                  tables[f"pw_float_{ratio}"])


def test_piecewise_survey_known_answers():
    # ratio 1.9 on [-13, 13]
    got = rel_index.piecewise_index(np.arange(-13, 14), 1.9, 3.8, 15.2)
    want = [-3] * 10 + [-2, -2, -1, 0, 1, 2, 2] + [3] * 10
    assert got.tolist() == want


METHODS = {"euc": rel_index.EUCLIDEAN, "quant": rel_index.QUANT, "product": rel_index.PRODUCT,
           "rows": rel_index.CROSS_ROWS, "cols": rel_index.CROSS_COLS}
GRIDS = [(14, 14, 1, 1.9), (7, 7, 0, 1.9), (5, 9, 2, 1.9), (14, 14, 1, 3.3), (24, 24, 1, 1.9)]


@pytest.mark.parametrize("mname", list(METHODS))
@pytest.mark.parametrize("h,w,skip,ratio", GRIDS)
def test_irpe_bucket_ids_bit_exact(tables, mname, h, w, skip, ratio):
    ids, nb = rel_index.irpe_bucket_ids(METHODS[mname], h, w, skip, 1 * ratio, 2 * ratio, 8 * ratio)
    key = f"{mname}_{h}_{w}_{skip}_{ratio}"
    assert nb == int(tables["nb_" + key])
    np.testing.assert_array_equal(ids, tables["ids_" + key])


def test_irpe_product_survey_known_answers():
    ids, nb = rel_index.irpe_bucket_ids(rel_index.PRODUCT, 14, 14, 1, 1.9, 3.8, 15.2)
    assert nb == 50 and ids.min() == 0 and ids.max() == 49 and len(np.unique(ids)) == 50
    assert ids[1, 1:16].tolist() == [24, 23, 22, 22, 21, 21, 21, 21, 21, 21, 21, 21, 21, 21, 17]
    assert (ids[0] == 49).all() and (ids[:, 0] == 49).all()


@pytest.mark.parametrize("grid", [14, 4, 7])
def test_autoformer_rel_index_bit_exact(tables, grid):
    iv, ih = rel_index.autoformer_rel_index(grid, 14)
    np.testing.assert_array_equal(iv, tables[f"af_idx_v_{grid}"])
    np.testing.assert_array_equal(ih, tables[f"af_idx_h_{grid}"])
    vals = set(np.unique(iv)) | set(np.unique(ih))
    assert 1 not in vals and 29 not in vals


def test_rpe_index_matches_reference_op(golden_dir):
    g = np.load(golden_dir / "rpe_index.npz")
    B, H, L, nb = 4, 3, 50, 50
    x = rand((B, H, L, nb), 5).numpy()
    idx = np.random.default_rng(6).integers(0, nb, (L, L)).astype(np.int32)
    gy = rand((B, H, L, L), 8).numpy()
    np.testing.assert_array_equal(rel_index.rpe_index_fwd(x, idx), g["y"])
    np.testing.assert_allclose(rel_index.rpe_index_bwd(gy, idx, nb), g["gx"], rtol=0, atol=1e-5)
    # plain-C restatement
    from oracle.build_ref import build_c_oracle
    lib = ctypes.CDLL(str(build_c_oracle()))
    y = np.empty((B, H, L, L), np.float32)
    P = ctypes.c_void_p
    lib.oracle_rpe_index_fwd_f32(P(x.ctypes.data), P(idx.ctypes.data), P(y.ctypes.data), B, H, L, L, nb)
    np.testing.assert_array_equal(y, g["y"])
    gx = np.zeros((B, H, L, nb), np.float32)
    lib.oracle_rpe_index_bwd_f32(P(gx.ctypes.data), P(gy.ctypes.data), P(idx.ctypes.data), B, H, L, L, nb)
    np.testing.assert_allclose(gx, g["gx"], rtol=0, atol=1e-5)


# ------------------------------------------------------------------------------------------
sys.path.insert(0, str(ROOT / "tests" / "golden"))
from make_golden import IRPE_CASES, MICRO_SPECS  # noqa: E402  (specs only; no reference import)


@pytest.mark.parametrize("name", list(MICRO_SPECS))
def test_supernet_oracle_matches_reference(golden_dir, name):
    g = np.load(golden_dir / "supernet_micro.npz")
    spec, batch, configs = MICRO_SPECS[name]
    images = rand((batch, 3, spec.img_size, spec.img_size), seed=11)
    targets = torch.from_numpy(np.random.default_rng(13).integers(0, spec.num_classes, batch))
    for ci, cfg in enumerate(configs):
        sd = {k: v.clone().requires_grad_(True) for k, v in vo.init_params(spec, seed=7).items()}
        logits = vo.supernet_forward(sd, cfg, images, spec)
        loss = torch.nn.functional.cross_entropy(logits, targets)
        loss.backward()
        key = f"{name}_c{ci}"
        np.testing.assert_allclose(logits.detach().numpy(), g[key + "_logits"], rtol=2e-4, atol=2e-5)
        assert abs(loss.item() - float(g[key + "_loss"])) < 1e-5
        assert vo.sampled_param_count(cfg, spec) == int(g[key + "_numel"])
        none = set(g[key + "_none"].tolist())
        for pn, p in sd.items():
            if pn in none:
                # reference: parameters of identity layers get no grad at all
                assert p.grad is None or float(p.grad.abs().sum()) == 0.0, pn
                assert int(pn.split(".")[1]) >= cfg["layer_num"]
            else:
                check_summary(g, f"{key}_grad_{pn}", p.grad, 2e-4, what=f"{key} grad {pn}")


def _irpe_params(name):
    """Recreate the seeded parameters make_golden gave the reference RPEAttention."""
    rpe_on, mode, shared, method, C, heads, grid = IRPE_CASES[name]
    g = np.load(ROOT / "tests" / "golden" / "irpe_attention.npz")
    shapes = {k[len(name) + 7:]: tuple(int(v) for v in g[k]) for k in g.files if k.startswith(name + "_shape_")}
    return shapes


@pytest.mark.parametrize("name", list(IRPE_CASES))
def test_irpe_attention_oracle_matches_reference(golden_dir, name):
    g = np.load(golden_dir / "irpe_attention.npz")
    rpe_on, mode, shared, method, C, heads, grid = IRPE_CASES[name]
    if method == "cross":
        ids = tuple(rel_index.irpe_bucket_ids(m, grid, grid, 1, 1.9, 3.8, 15.2)[0]
                    for m in (rel_index.CROSS_ROWS, rel_index.CROSS_COLS))
    else:
        mid = {"product": rel_index.PRODUCT, "euc": rel_index.EUCLIDEAN, "quant": rel_index.QUANT}[method]
        ids, nb = rel_index.irpe_bucket_ids(mid, grid, grid, 1, 1.9, 3.8, 15.2)
    N, B = grid * grid + 1, 2
    # parameter order of RPEAttention.named_parameters(): qkv.weight, qkv.bias, proj.weight,
    # proj.bias, then rpe_q / rpe_k / rpe_v lookup tables in attribute order
    shapes = _irpe_params(name)
    params, seed = {}, 100
    for pn, shape in shapes.items():
        seed += 1
        params[pn] = rand(shape, seed, 0.3 if "lookup" in pn else 0.08).requires_grad_(True)
    def tab(w):
        hits = [v for k, v in params.items() if k.startswith(f"rpe_{w}.")]   # cross: rp_rows, rp_cols
        return None if not hits else (hits[0] if len(hits) == 1 else tuple(hits))
    x = rand((B, N, C), 99).requires_grad_(True)
    gy = rand((B, N, C), 98)
    y = vo.rpe_attention(x, params["qkv.weight"], params["qkv.bias"], params["proj.weight"], params["proj.bias"],
                         heads, ids, rpe_q=tab("q"), rpe_k=tab("k"), rpe_v=tab("v"),
                         mode="bias" if mode == "bias" else "contextual")
    y.backward(gy)
    check_summary(g, f"{name}_y", y, 1e-5)
    check_summary(g, f"{name}_gx", x.grad, 1e-4)
    for pn, p in params.items():
        check_summary(g, f"{name}_grad_{pn}", p.grad, 1e-4, what=f"{name} grad {pn}")


@pytest.mark.parametrize("L,B,heads,causal", [(50, 3, 2, False), (77, 2, 2, True)])
def test_clip_attention_oracle_matches_torch_mha(L, B, heads, causal):
    """The TinyCLIP attention restatement against nn.MultiheadAttention — the call the reference's
    default branch makes (open_clip/model.py:246-248) — forward and input gradient."""
    E = 64 * heads
    torch.manual_seed(5)
    mha = torch.nn.MultiheadAttention(E, heads)
    x = torch.randn(L, B, E, requires_grad=True)
    mask = torch.full((L, L), float("-inf")).triu_(1) if causal else None
    ref = mha(x, x, x, need_weights=False, attn_mask=mask)[0]
    gy = torch.randn_like(ref)
    ref.backward(gy)
    x2 = x.detach().clone().requires_grad_(True)
    y = vo.clip_attention(x2, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias,
                          heads, mask)
    y.backward(gy)
    assert torch.allclose(y, ref, atol=2e-5, rtol=1e-4)
    assert torch.allclose(x2.grad, x.grad, atol=2e-5, rtol=1e-4)


from make_golden import TINYVIT_CASES  # noqa: E402


def _tinyvit_params(g, name):
    shapes = {k[len(name) + 7:]: tuple(int(v) for v in g[k]) for k in g.files if k.startswith(name + "_shape_")}
    params, seed = {}, 300
    for pn, shape in shapes.items():
        seed += 1
        if pn == "norm.weight":
            params[pn] = (1.0 + rand(shape, seed, 0.1)).requires_grad_(True)
        else:
            params[pn] = rand(shape, seed, 0.5 if "attention_biases" in pn else 0.08).requires_grad_(True)
    return params


@pytest.mark.parametrize("name", list(TINYVIT_CASES))
def test_tinyvit_attention_oracle_matches_reference(golden_dir, name):
    g = np.load(golden_dir / "tinyvit_attention.npz")
    dim, key_dim, heads, ratio, res, B = TINYVIT_CASES[name]
    idxs, n_off = vo.tinyvit_bias_idxs(res)
    np.testing.assert_array_equal(idxs.numpy(), g[f"{name}_idxs"])           # integer table: bit exact
    P = _tinyvit_params(g, name)
    assert P["attention_biases"].shape == (heads, n_off)
    N = res[0] * res[1]
    x = rand((B, N, dim), 299).requires_grad_(True)
    y = vo.tinyvit_attention(x, P, heads, key_dim, int(ratio * key_dim), idxs)
    y.backward(rand((B, N, dim), 298))
    check_summary(g, f"{name}_y", y, 1e-5)
    check_summary(g, f"{name}_gx", x.grad, 1e-4)
    for pn, p in P.items():
        check_summary(g, f"{name}_grad_{pn}", p.grad, 1e-4, what=f"{name} grad {pn}")


@pytest.mark.skipif(not Path("/root/reference").exists(), reason="reference checkout only exists in the build container")
def test_irpe_bucket_ids_property_against_the_reference():
    """Beyond the committed fixtures: random (method, grid, skip, ratio) draws, the numpy restatement
    against the reference's own get_bucket_ids_2d (irpe.py:364-415) imported in place — bit exact."""
    from hypothesis import given, settings, strategies as st
    import make_golden as mg
    mg.install_shims()
    sys.path.insert(0, "/root/reference/iRPE/DeiT-with-iRPE")
    sys.dont_write_bytecode = True
    import irpe

    @settings(max_examples=120, deadline=None)
    @given(mid=st.sampled_from([0, 1, 3, 41, 42]), h=st.integers(1, 15), w=st.integers(1, 15),
           skip=st.integers(0, 2), ratio=st.floats(1.0, 4.0, allow_nan=False, width=32))
    def check(mid, h, w, skip, ratio):
        irpe.BUCKET_IDS_BUF.clear()
        want, nb_want = irpe.get_bucket_ids_2d(mid, h, w, skip, 1 * ratio, 2 * ratio, 8 * ratio)
        ids, nb = rel_index.irpe_bucket_ids(mid, h, w, skip, 1 * ratio, 2 * ratio, 8 * ratio)
        assert nb == int(nb_want)
        np.testing.assert_array_equal(ids, want.numpy())

    check()


@pytest.mark.skipif(not Path("/root/reference").exists(), reason="reference checkout only exists in the build container")
def test_autoformer_rel_index_property_against_the_reference():
    """Random (grid, max_relative_position) draws — including clamps that bind — against the reference's
    RelativePosition2D_super (multihead_super.py:40-66) imported in place: its tables are set to
    arange so the returned embedding decodes to (idx_v, idx_h)."""
    import importlib.util
    from hypothesis import given, settings, strategies as st
    import make_golden as mg
    mg.install_shims()
    sys.dont_write_bytecode = True
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "model" or k.startswith("model.")}
    sys.path.insert(0, "/root/reference/AutoFormer")
    try:
        for k in saved:
            sys.modules.pop(k, None)
        from model.module.multihead_super import RelativePosition2D_super  # reference, unmodified

        @settings(max_examples=80, deadline=None)
        @given(grid=st.integers(1, 15), max_rel=st.integers(1, 15))
        def check(grid, max_rel):
            n = grid * grid + 1
            m = RelativePosition2D_super(1, max_rel)
            rows = 2 * max_rel + 2
            with torch.no_grad():
                m.embeddings_table_v.copy_(torch.arange(rows, dtype=torch.float32).view(rows, 1))
                m.embeddings_table_h.copy_(1000.0 * torch.arange(rows, dtype=torch.float32).view(rows, 1))
            m.set_sample_config(1)
            code = m(n, n)[..., 0].round().long()
            iv, ih = rel_index.autoformer_rel_index(grid, max_rel)
            np.testing.assert_array_equal(iv, (code % 1000).numpy())
            np.testing.assert_array_equal(ih, (code // 1000).numpy())

        check()
    finally:
        sys.path.remove("/root/reference/AutoFormer")
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


@pytest.mark.skipif(not Path("/root/reference").exists(), reason="reference checkout only exists in the build container")
@pytest.mark.parametrize("rpe_on,mode,method,shared", [
    ("k", "ctx", "euc", True), ("qk", "ctx", "quant", False), ("qkv", "ctx", "product", True),
    ("kv", "ctx", "cross", False), ("v", "ctx", "product", False), ("q", "ctx", "cross", True),
    ("qk", "bias", "product", False), ("k", "bias", "cross", True), ("q", "bias", "euc", False),
])
def test_rpe_attention_oracle_against_the_reference_module(rpe_on, mode, method, shared):
    """A sweep beyond the committed fixtures: the oracle's RPEAttention restatement against the
    reference's own module (rpe_vision_transformer.py:45-97, timm stubbed) on small grids, every
    rpe_on / mode / method / head-sharing family, forward and all gradients."""
    import make_golden as mg
    mg.install_shims()
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference/iRPE/DeiT-with-iRPE")
    # the reference must take its pure-PyTorch gather (irpe.py:8-15 falls back on ImportError), not a
    # drop-in rpe_ops another test may have registered: import it fresh with rpe_ops blocked
    names = ("irpe", "rpe_vision_transformer", "rpe_ops", "rpe_ops.rpe_index")
    held = {k: sys.modules.pop(k, None) for k in names}
    sys.modules["rpe_ops"] = None
    sys.modules["rpe_ops.rpe_index"] = None
    try:
        import irpe
        from rpe_vision_transformer import RPEAttention
        irpe.BUCKET_IDS_BUF.clear()
        C, heads, grid, B = 32, 2, 4, 2
        cfg = irpe.get_rpe_config(ratio=1.9, method=method, mode=mode, shared_head=shared, skip=1, rpe_on=rpe_on)
        attn = RPEAttention(C, num_heads=heads, qkv_bias=True, rpe_config=cfg)
        seed = 700
        with torch.no_grad():
            for pn, p in attn.named_parameters():
                seed += 1
                p.copy_(rand(tuple(p.shape), seed, 0.3 if "lookup" in pn else 0.1))
        N = grid * grid + 1
        x = rand((B, N, C), 699).requires_grad_(True)
        gy = rand((B, N, C), 698)
        y = attn(x)
        y.backward(gy)
        P = {k: v.detach().clone().requires_grad_(True) for k, v in attn.named_parameters()}
        if method == "cross":
            ids = tuple(rel_index.irpe_bucket_ids(m, grid, grid, 1, 1.9, 3.8, 15.2)[0]
                        for m in (rel_index.CROSS_ROWS, rel_index.CROSS_COLS))
        else:
            mid = {"product": rel_index.PRODUCT, "euc": rel_index.EUCLIDEAN, "quant": rel_index.QUANT}[method]
            ids = rel_index.irpe_bucket_ids(mid, grid, grid, 1, 1.9, 3.8, 15.2)[0]

        def tab(w):
            hits = [v for k, v in P.items() if k.startswith(f"rpe_{w}.")]
            return None if not hits else (hits[0] if len(hits) == 1 else tuple(hits))
        x2 = x.detach().clone().requires_grad_(True)
        y2 = vo.rpe_attention(x2, P["qkv.weight"], P["qkv.bias"], P["proj.weight"], P["proj.bias"], heads, ids,
                              rpe_q=tab("q"), rpe_k=tab("k"), rpe_v=tab("v"),
                              mode="bias" if mode == "bias" else "contextual")
        y2.backward(gy)
        assert torch.allclose(y2, y, atol=1e-5, rtol=1e-4)
        assert torch.allclose(x2.grad, x.grad, atol=1e-5, rtol=1e-4)
        for k, p in attn.named_parameters():
            assert torch.allclose(P[k].grad, p.grad, atol=1e-5, rtol=1e-4), k
    finally:
        sys.path.remove("/root/reference/iRPE/DeiT-with-iRPE")
        for k in names:
            sys.modules.pop(k, None)
        sys.modules.update({k: v for k, v in held.items() if v is not None})


@pytest.mark.skipif(not Path("/root/reference").exists(), reason="reference checkout only exists in the build container")
def test_supernet_t_oracle_against_the_reference_model():
    """The shipped supernet-T geometry (BASELINE config 1 family) with two subnets drawn by the
    reference's own sampling rule: oracle logits and a few gradients against the reference model
    imported in place (supernet_transformer.py, torch._six shimmed)."""
    import random
    import make_golden as mg
    mg.install_shims()
    sys.dont_write_bytecode = True
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "model" or k.startswith("model.")}
    sys.path.insert(0, "/root/reference/AutoFormer")
    try:
        for k in saved:
            sys.modules.pop(k, None)
        from model.supernet_transformer import Vision_TransformerSuper  # reference, unmodified
        spec = vo.SUPERNET_T
        net = Vision_TransformerSuper(img_size=224, patch_size=16, embed_dim=spec.embed_dim, depth=spec.depth,
                                      num_heads=spec.num_heads, mlp_ratio=spec.mlp_ratio, qkv_bias=True, drop_rate=0.0,
                                      drop_path_rate=0.0, gp=True, num_classes=1000, max_relative_position=14,
                                      relative_position=True, change_qkv=True, abs_pos=True)
        sd = vo.init_params(spec, seed=3)
        net.load_state_dict(sd)
        net.train()
        rnd = random.Random(1)
        images = rand((1, 3, 224, 224), seed=41)
        for _ in range(2):
            cfg = vo.sample_configs(vo.SEARCH_SPACE["T"], rnd)
            net.zero_grad(set_to_none=True)
            net.set_sample_config(cfg)
            ref = net(images)
            ref.sum().backward()
            P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            out = vo.supernet_forward(P, cfg, images, spec)
            out.sum().backward()
            assert torch.allclose(out, ref, atol=2e-4, rtol=1e-3)
            for name in ("cls_token", "blocks.0.attn.qkv.weight", "blocks.0.attn.rel_pos_embed_k.embeddings_table_v",
                         "blocks.1.fc1.weight", "head.weight"):
                assert torch.allclose(P[name].grad, dict(net.named_parameters())[name].grad, atol=2e-4, rtol=1e-3), name
    finally:
        sys.path.remove("/root/reference/AutoFormer")
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


def test_clip_attention_oracle_is_pinned_against_the_reference_block():
    """vo.clip_attention vs the reference's OWN `ResidualAttentionBlock.attention`
    (TinyCLIP/src/open_clip/model.py:238-283, loaded unmodified through oracle/refload.py): the
    nn.MultiheadAttention branch (with and without the causal mask of the text tower, model.py:756-762) and the
    naive branch with the pruning multipliers head_z / hidden_z."""
    from oracle import refload
    if not refload.available():
        pytest.skip("reference not available (build container / staged baseline/_ref only)")
    m = refload.open_clip_model()
    torch.manual_seed(0)
    for d_model, heads, length, batch in ((128, 2, 50, 3), (512, 8, 77, 2)):
        blk = m.ResidualAttentionBlock(d_model=d_model, n_head=heads)
        a = blk.attn
        x = torch.randn(length, batch, d_model)
        causal = torch.full((length, length), float("-inf")).triu_(1)
        cases = [(None, None, None), (causal, None, None), (None, torch.rand(heads), torch.rand(d_model))]
        for mask, hz, hd in cases:
            with torch.no_grad():
                want = blk.attention(x, attn_mask=mask, head_z=hz, hidden_z=hd)
                got = vo.clip_attention(x, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias, heads,
                                        attn_mask=mask, head_z=hz, hidden_z=hd)
            assert float((want - got).abs().max()) < 2e-6
