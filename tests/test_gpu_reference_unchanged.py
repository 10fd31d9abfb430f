"""The reference's OWN files, unchanged, on the GPU over cream_b200 (BASELINE.json north_star: "so
AutoFormer/model/supernet_transformer.py and iRPE/DeiT-with-iRPE/rpe_ops drop in unchanged").

Needs the reference: `$CREAM_REFERENCE`, /root/reference, or baseline/_ref staged by
scripts/stage_reference.py (git-ignored, travels to the GPU box).  Skipped when absent.
"""
import random
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from oracle import refload, rel_index, vit_oracle as vo  # noqa: E402
from tests.helpers import rand, rel_err  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refload.available(), reason="reference not staged")]


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cream_b200 import _lib, ops
    _lib.load()
    ops.SHADOWS.clear()
    yield
    torch.cuda.synchronize()


def _ctor(spec):
    return dict(img_size=spec.img_size, patch_size=spec.patch_size, embed_dim=spec.embed_dim, depth=spec.depth,
                num_heads=spec.num_heads, mlp_ratio=spec.mlp_ratio, qkv_bias=True, drop_rate=0.0, drop_path_rate=0.0,
                gp=True, num_classes=spec.num_classes, max_relative_position=14, relative_position=True,
                change_qkv=True, abs_pos=True)


def _run(net, cfg, images, targets):
    net.zero_grad(set_to_none=True)
    net.set_sample_config(cfg)
    logits = net(images)
    loss = F.cross_entropy(logits.float(), targets)
    loss.backward()
    return logits.detach().float().cpu(), {k: (None if p.grad is None else p.grad.float().cpu()) for k, p in net.named_parameters()}


# ------------------------------------------------------------------------------------------------
# (a) AutoFormer/model/supernet_transformer.py forward + backward over the drop-in modules
# ------------------------------------------------------------------------------------------------
def test_reference_supernet_file_runs_forward_backward_over_the_drop_ins():
    from cream_b200.autoformer.model.supernet_transformer import Vision_TransformerSuper as Mirror
    spec = vo.SUPERNET_T
    cfg = {'layer_num': 12, 'embed_dim': [192] * 12, 'num_heads': [3] * 12, 'mlp_ratio': [3.5] * 12}   # BASELINE config 1
    sd = vo.init_params(spec, seed=7)
    images = rand((4, 3, 224, 224), seed=11).cuda()
    targets = torch.from_numpy(np.random.default_rng(13).integers(0, 1000, 4)).cuda()
    ref_mod = refload.autoformer("cream")            # the reference file, cream_b200 modules underneath
    assert ref_mod.__file__.endswith("AutoFormer/model/supernet_transformer.py")
    net = ref_mod.Vision_TransformerSuper(**_ctor(spec))
    assert type(net.blocks[0].fc1).__module__.startswith("cream_b200.")
    assert type(net.blocks[0]).__module__ == ref_mod.__name__, "blocks are the reference's TransformerEncoderLayer"
    net.load_state_dict(sd)
    net = net.cuda().train()
    y_mod, g_mod = _run(net, cfg, images, targets)
    # fp32 oracle
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = vo.supernet_forward(P, cfg, images.cpu(), spec)
    F.cross_entropy(ref, targets.cpu()).backward()
    e = rel_err(y_mod, ref.detach())
    assert e < 1e-2, f"reference file over drop-ins vs oracle: logits {e:.3e}"
    worst = max(rel_err(g_mod[k], v.grad) for k, v in P.items() if v.grad is not None)
    assert worst < 4e-2, f"worst grad {worst:.3e}"
    assert all(g_mod[k] is None for k, v in P.items() if v.grad is None)
    # the fused mirror on the same weights
    mirror = Mirror(**_ctor(spec))
    mirror.load_state_dict(sd)
    y_fused, g_fused = _run(mirror.cuda().train(), cfg, images, targets)
    assert rel_err(y_fused, y_mod) < 1e-2
    print(f"\n[reference file over drop-ins] logits vs oracle {e:.3e}, worst grad {worst:.3e}, "
          f"vs fused mirror {rel_err(y_fused, y_mod):.3e}")


def test_fuse_reference_patches_the_reference_class_itself():
    """`fuse_reference` routes the forward of the reference's own class - here over the reference's
    OWN modules, nothing of cream_b200 in the module tree - through the fused engine; eval falls back
    to nothing: the same engine runs without saving activations."""
    from cream_b200.autoformer.model.supernet_transformer import fuse_reference
    spec, batch, cfgs = vo.SUPERNET_T, 4, None
    sd = vo.init_params(spec, seed=7)
    ref_mod = refload.autoformer("reference")
    cls = type("FusedReferenceSupernet", (ref_mod.Vision_TransformerSuper,), {})     # keep the imported class pristine
    fuse_reference(cls)
    net = cls(**_ctor(spec))
    assert type(net.blocks[0].fc1).__module__ == "model.module.Linear_super"
    net.load_state_dict(sd)
    net = net.cuda().train()
    images = rand((batch, 3, 224, 224), seed=11).cuda()
    targets = torch.from_numpy(np.random.default_rng(13).integers(0, 1000, batch)).cuda()
    rnd = random.Random(5)
    for _ in range(2):
        cfg = vo.sample_configs(vo.SEARCH_SPACE["T"], rnd)
        y, g = _run(net, cfg, images, targets)
        P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        ref = vo.supernet_forward(P, cfg, images.cpu(), spec)
        F.cross_entropy(ref, targets.cpu()).backward()
        assert rel_err(y, ref.detach()) < 1e-2
        for k, v in P.items():
            if v.grad is None:
                assert g[k] is None, k
            else:
                assert rel_err(g[k], v.grad) < 4e-2, k
    net.eval()
    with torch.no_grad():
        ye = net(images)
    assert rel_err(ye.float().cpu(), y) < 1e-3
    assert net.get_sampled_params_numel(cfg) == vo.sampled_param_count(cfg, spec)


# ------------------------------------------------------------------------------------------------
# (b) irpe.py picks up cream_b200/rpe_ops and runs on CUDA
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rpe_on,method,mode,shared", [("k", "product", "ctx", True), ("qkv", "product", "ctx", False),
                                                        ("k", "cross", "ctx", True), ("qk", "euc", "bias", False)])
def test_reference_irpe_runs_on_cream_rpe_ops(rpe_on, method, mode, shared):
    irpe_c, irpe_r = refload.irpe("cream"), refload.irpe("reference")
    from cream_b200.rpe_ops.rpe_index import RPEIndexFunction
    assert irpe_c.RPEIndexFunction is RPEIndexFunction and irpe_r.RPEIndexFunction is None
    B, H, L, D = 4, 6, 197, 64
    cfg = irpe_c.get_rpe_config(ratio=1.9, method=method, mode=mode, shared_head=shared, skip=1, rpe_on=rpe_on)
    mods_c = irpe_c.build_rpe(cfg, head_dim=D, num_heads=H)
    mods_r = irpe_r.build_rpe(irpe_r.get_rpe_config(ratio=1.9, method=method, mode=mode, shared_head=shared, skip=1,
                                                     rpe_on=rpe_on), head_dim=D, num_heads=H)
    x = rand((B, H, L, D), 71).cuda()
    attn = torch.softmax(rand((B, H, L, L), 72).cuda(), -1)
    for which, mc, mr in zip("qkv", mods_c, mods_r):
        if mc is None:
            assert mr is None
            continue
        mc, mr = mc.cuda(), mr.cuda()
        seed = 80
        with torch.no_grad():
            for (n, p), (_, q) in zip(mc.named_parameters(), mr.named_parameters()):
                seed += 1
                p.copy_(rand(tuple(p.shape), seed, 0.2))
                q.copy_(p)
        inp = attn if which == "v" else x
        a = inp.clone().requires_grad_(True)
        b = inp.clone().requires_grad_(True)
        yc, yr = mc(a), mr(b)
        if mode == "ctx" and which != "v":   # int32 ids are selected exactly when the native op is present (irpe.py:563-565)
            sub_c, sub_r = (mc.rp_rows, mr.rp_rows) if method == "cross" else (mc, mr)
            assert sub_c._rp_bucket_buf[1].dtype == torch.int32 and sub_r._rp_bucket_buf[1].dtype == torch.int64
        assert torch.equal(yc, yr), f"rpe_{which}: forward through cream rpe_ops must be bit-identical to the fallback gather"
        gy = torch.randn_like(yc)
        yc.backward(gy)
        yr.backward(gy)
        if a.grad is not None:
            assert rel_err(a.grad, b.grad) < 1e-5
        for (n, p), (_, q) in zip(mc.named_parameters(), mr.named_parameters()):
            assert rel_err(p.grad, q.grad) < 1e-5, n


def test_reference_rpe_attention_over_cream_rpe_ops_matches_fused_module():
    """rpe_vision_transformer.RPEAttention (reference, unchanged) with cream rpe_ops underneath, against
    cream_b200's fused RPEAttention on the same weights - the two ways a DeiT+iRPE user can adopt the
    library (swap rpe_ops only / swap the attention module)."""
    from cream_b200.irpe_attention import RPEAttention
    vit = refload.rpe_vision_transformer("cream")
    C, heads, N, B = 384, 6, 197, 4
    cfg = vit.irpe.get_rpe_config(ratio=1.9, method="product", mode="ctx", shared_head=True, skip=1, rpe_on="k")
    ref_attn = vit.RPEAttention(C, num_heads=heads, qkv_bias=True, rpe_config=cfg).cuda()
    ours = RPEAttention(C, num_heads=heads, qkv_bias=True, rpe_on="k", method="product", mode="ctx", shared_head=True).cuda()
    seed = 600
    with torch.no_grad():
        for (n, p) in ref_attn.named_parameters():
            seed += 1
            p.copy_(rand(tuple(p.shape), seed, 0.02 if "lookup" in n else 0.05))
        ours.load_state_dict(ref_attn.state_dict())
    x = rand((B, N, C), 599).cuda()
    a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ref_attn(a), ours(b)
    gy = rand((B, N, C), 598).cuda()
    ya.backward(gy)
    yb.backward(gy.to(yb.dtype))
    assert rel_err(yb.float(), ya) < 1e-2
    assert rel_err(b.grad.float(), a.grad) < 2e-2
    for (n, p), (_, q) in zip(ref_attn.named_parameters(), ours.named_parameters()):
        assert rel_err(q.grad.float(), p.grad) < 2e-2, n


# ------------------------------------------------------------------------------------------------
# (c) supernet_engine.train_one_epoch / evaluate drive the fused model
# ------------------------------------------------------------------------------------------------
def test_reference_train_one_epoch_and_evaluate_over_the_fused_model():
    from cream_b200.autoformer.model.supernet_transformer import fuse_reference
    eng = refload.supernet_engine()
    ref_mod = refload.autoformer("cream")
    cls = type("FusedSupernetOverDropIns", (ref_mod.Vision_TransformerSuper,), {})
    fuse_reference(cls)
    spec, space = vo.SUPERNET_T, vo.SEARCH_SPACE["T"]
    sd = vo.init_params(spec, seed=3)
    net = cls(**_ctor(spec))
    net.load_state_dict(sd)
    net = net.cuda()
    steps, batch, lr = 3, 8, 1e-3
    data = [(rand((batch, 3, 224, 224), seed=40 + s), torch.from_numpy(np.random.default_rng(50 + s).integers(0, 1000, batch)))
            for s in range(steps)]
    opt = torch.optim.AdamW(net.parameters(), lr=lr, weight_decay=0.05)
    stats = eng.train_one_epoch(net, torch.nn.CrossEntropyLoss(), data, opt, torch.device("cuda"), epoch=0,
                                loss_scaler=None, max_norm=0, model_ema=None, mixup_fn=None, amp=False,
                                choices=space, mode='super')
    # the same loop on the fp32 oracle: train_one_epoch seeds `random` with the epoch (supernet_engine.py:36)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ropt = torch.optim.AdamW(list(P.values()), lr=lr, weight_decay=0.05)
    rnd, losses = random.Random(0), []
    for images, targets in data:
        cfg = vo.sample_configs(space, rnd)
        ropt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(vo.supernet_forward(P, cfg, images, spec), targets)
        loss.backward()
        ropt.step()
        losses.append(float(loss))
    assert abs(stats["loss"] - sum(losses) / len(losses)) < 2e-2, (stats, losses)
    # evaluate(): samples a config, prints the sampled parameter count, runs the eval forward
    random.seed(1)
    out = eng.evaluate(data[:1], net, torch.device("cuda"), amp=False, choices=space, mode='super')
    assert np.isfinite(out["loss"]) and 0.0 <= out["acc1"] <= 100.0
    # amp=True path: autocast around the model, a timm-NativeScaler-shaped loss scaler
    class Scaler:
        def __init__(self):
            self._s = torch.amp.GradScaler("cuda")
        def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False):
            self._s.scale(loss).backward(create_graph=create_graph)
            self._s.step(optimizer)
            self._s.update()
    stats2 = eng.train_one_epoch(net, torch.nn.CrossEntropyLoss(), data, opt, torch.device("cuda"), epoch=1,
                                 loss_scaler=Scaler(), max_norm=0, amp=True, choices=space, mode='super')
    assert np.isfinite(stats2["loss"])
