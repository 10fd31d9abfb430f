"""GPU parity on BASELINE.json's OWN configurations, against the live oracle (fp32, CPU) and with a
same-precision comparator that justifies every tolerance:

    err(cream_b200 bf16 path, fp32 oracle)  <=  RATIO * err(plain torch bf16-autocast path, fp32 oracle)

The comparator is the reference model itself under `torch.autocast(bfloat16)` on the GPU when the
reference is staged (baseline/_ref), else the oracle restatement under the same autocast: that is the
precision the reference trains at (it uses fp16 autocast, supernet_engine.py:68).  A kernel whose
error against fp32 is no larger than the stock mixed-precision path's is "within bf16" by measurement,
not by prose.
"""
import os
import random
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent

from oracle import refload, rel_index, vit_oracle as vo  # noqa: E402
from tests.helpers import rand, rel_err  # noqa: E402

pytestmark = pytest.mark.gpu

RATIO = 1.5          # cream error may be at most this multiple of the comparator's error ...
FLOOR = 1.0e-3       # ... or this absolute relative-L2 error (BASELINE.json's 1e-3), whichever is larger


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cream_b200 import _lib, ops
    _lib.load()
    ops.SHADOWS.clear()
    yield
    torch.cuda.synchronize()


def _mirror(spec, drop_path=0.0):
    from cream_b200.autoformer.model.supernet_transformer import Vision_TransformerSuper
    return Vision_TransformerSuper(img_size=spec.img_size, patch_size=spec.patch_size, embed_dim=spec.embed_dim,
                                   depth=spec.depth, num_heads=spec.num_heads, mlp_ratio=spec.mlp_ratio, qkv_bias=True,
                                   drop_rate=0.0, drop_path_rate=drop_path, gp=True, num_classes=spec.num_classes,
                                   max_relative_position=14, relative_position=True, change_qkv=True, abs_pos=True)


def _oracle_step(sd, cfg, images, targets, spec):
    """fp32 CPU oracle: logits, loss, {name: grad or None}."""
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits = vo.supernet_forward(P, cfg, images, spec)
    loss = F.cross_entropy(logits, targets)
    loss.backward()
    return logits.detach(), float(loss), {k: (None if v.grad is None else v.grad) for k, v in P.items()}


def _comparator_step(sd, cfg, images, targets, spec):
    """Stock mixed precision on the GPU: the reference model (when staged) or the oracle restatement
    under torch.autocast(bfloat16).  Returns (logits fp32 cpu, {name: grad cpu})."""
    dev = "cuda"
    if refload.available():
        mod = refload.autoformer("reference")
        net = mod.Vision_TransformerSuper(img_size=spec.img_size, patch_size=spec.patch_size, embed_dim=spec.embed_dim,
                                          depth=spec.depth, num_heads=spec.num_heads, mlp_ratio=spec.mlp_ratio,
                                          qkv_bias=True, drop_rate=0.0, drop_path_rate=0.0, gp=True,
                                          num_classes=spec.num_classes, max_relative_position=14,
                                          relative_position=True, change_qkv=True, abs_pos=True)
        net.load_state_dict(sd)
        net = net.to(dev).train()
        net.set_sample_config(cfg)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = net(images.to(dev))
            loss = F.cross_entropy(logits.float(), targets.to(dev))
        loss.backward()
        grads = {k: (None if p.grad is None else p.grad.float().cpu()) for k, p in net.named_parameters()}
        return logits.detach().float().cpu(), grads, "reference model, autocast(bf16)"
    P = {k: v.to(dev).requires_grad_(True) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = vo.supernet_forward(P, cfg, images.to(dev), spec)
        loss = F.cross_entropy(logits.float(), targets.to(dev))
    loss.backward()
    return logits.detach().float().cpu(), {k: (None if v.grad is None else v.grad.float().cpu()) for k, v in P.items()}, \
        "oracle restatement, autocast(bf16)"


def _cream_step(sd, cfg, images, targets, spec):
    from cream_b200 import ops
    ops.SHADOWS.clear()
    net = _mirror(spec)
    net.load_state_dict(sd)
    net = net.cuda().train()
    net.set_sample_config(cfg)
    logits = net(images.cuda())
    loss = F.cross_entropy(logits.float(), targets.cuda())
    loss.backward()
    grads = {k: (None if p.grad is None else p.grad.float().cpu()) for k, p in net.named_parameters()}
    return logits.detach().float().cpu(), float(loss), grads


def _compare(tag, spec, cfg, batch, seed=11):
    sd = vo.init_params(spec, seed=7)
    images = rand((batch, 3, spec.img_size, spec.img_size), seed=seed)
    targets = torch.from_numpy(np.random.default_rng(13).integers(0, spec.num_classes, batch))
    ref_logits, ref_loss, ref_grads = _oracle_step(sd, cfg, images, targets, spec)
    cmp_logits, cmp_grads, cmp_name = _comparator_step(sd, cfg, images, targets, spec)
    our_logits, our_loss, our_grads = _cream_step(sd, cfg, images, targets, spec)
    e_our, e_cmp = rel_err(our_logits, ref_logits), rel_err(cmp_logits, ref_logits)
    print(f"\n[{tag}] comparator = {cmp_name}")
    print(f"[{tag}] logits rel-L2 vs fp32 oracle: cream {e_our:.3e}   comparator {e_cmp:.3e}")
    assert e_our <= max(RATIO * e_cmp, FLOOR), f"{tag}: logits {e_our:.3e} vs comparator {e_cmp:.3e}"
    assert abs(our_loss - ref_loss) <= max(2e-3 * abs(ref_loss), 2e-3)
    none = [k for k, g in ref_grads.items() if g is None]
    worst = (0.0, 0.0, "")
    tot_our = tot_cmp = tot_ref = 0.0
    for k, g in ref_grads.items():
        if g is None:
            assert our_grads[k] is None, f"{tag}: {k} belongs to an identity layer and must get no gradient"
            continue
        assert our_grads[k] is not None, k
        a, c = rel_err(our_grads[k], g), rel_err(cmp_grads[k], g)
        tot_our += float((our_grads[k].double() - g.double()).pow(2).sum())
        tot_cmp += float((cmp_grads[k].double() - g.double()).pow(2).sum())
        tot_ref += float(g.double().pow(2).sum())
        if a > worst[0]:
            worst = (a, c, k)
        # per parameter: never worse than the stock mixed-precision path by more than RATIO (small
        # tensors - biases, tables - are noisier: allow the bf16 epsilon 2^-8 as a floor)
        assert a <= max(RATIO * c, 2.0 ** -8), f"{tag}: grad {k} {a:.3e} vs comparator {c:.3e}"
    g_our, g_cmp = (tot_our / tot_ref) ** 0.5, (tot_cmp / tot_ref) ** 0.5
    print(f"[{tag}] all gradients, global rel-L2: cream {g_our:.3e}   comparator {g_cmp:.3e};  worst parameter "
          f"{worst[2]}: cream {worst[0]:.3e} comparator {worst[1]:.3e};  {len(none)} parameters without grad")
    assert g_our <= max(RATIO * g_cmp, FLOOR)
    return none


# ------------------------------------------------------------------------------------------------
# (i) BASELINE config 1 exactly: supernet-T, fixed sample (192, 3 heads, 12 layers), bs 4, 224^2
# ------------------------------------------------------------------------------------------------
def test_config1_supernet_t_fixed_sample_bs4():
    cfg = {'layer_num': 12, 'embed_dim': [192] * 12, 'num_heads': [3] * 12, 'mlp_ratio': [3.5] * 12}
    none = _compare("c1", vo.SUPERNET_T, cfg, batch=4)
    assert len(none) == 32, "32 of 232 parameters (blocks 12, 13) must keep grad None (SURVEY.md 8c)"


# ------------------------------------------------------------------------------------------------
# (ii) supernet-S largest / smallest subnet, supernet-B subnet, bs 8
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,spec,cfg", [
    ("c3-max", vo.SUPERNET_S, {'layer_num': 14, 'embed_dim': [448] * 14, 'num_heads': [7] * 14, 'mlp_ratio': [4.0] * 14}),
    ("c3-min", vo.SUPERNET_S, {'layer_num': 12, 'embed_dim': [320] * 12, 'num_heads': [5] * 12, 'mlp_ratio': [3.0] * 12}),
    ("c3-mixed", vo.SUPERNET_S, {'layer_num': 13, 'embed_dim': [384] * 13, 'num_heads': [5, 6, 7, 6, 5, 7, 7, 5, 6, 6, 7, 5, 6],
                                 'mlp_ratio': [3.0, 3.5, 4.0, 4.0, 3.5, 3.0, 3.5, 4.0, 3.0, 3.5, 4.0, 3.0, 3.5]}),
    ("c5-max", vo.SUPERNET_B, {'layer_num': 16, 'embed_dim': [624] * 16, 'num_heads': [10] * 16, 'mlp_ratio': [4.0] * 16}),
    ("c5-mixed", vo.SUPERNET_B, {'layer_num': 14, 'embed_dim': [528] * 14, 'num_heads': [9, 10] * 7,
                                 'mlp_ratio': [3.0, 3.5, 4.0, 3.5, 3.0, 4.0, 3.5] * 2}),
])
def test_config3_config5_subnets_bs8(tag, spec, cfg):
    _compare(tag, spec, cfg, batch=8)


# ------------------------------------------------------------------------------------------------
# (iii) BASELINE config 2 block: DeiT-S attention with iRPE (product, contextual, shared head, keys)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rpe_on", ["k", "qkv"])
def test_config2_deit_s_irpe_attention_block(rpe_on):
    from cream_b200.irpe_attention import RPEAttention
    C, heads, N, B = 384, 6, 197, 8
    m = RPEAttention(C, num_heads=heads, qkv_bias=True, rpe_on=rpe_on, method="product", mode="ctx",
                     shared_head=True).cuda()
    seed = 500
    with torch.no_grad():
        for pn, p in m.named_parameters():
            seed += 1
            p.copy_(rand(tuple(p.shape), seed, 0.02 if "lookup" in pn else 0.05))   # tables randn * 0.02
    x = rand((B, N, C), 499)
    gy = rand((B, N, C), 498)
    ids, nb = rel_index.irpe_bucket_ids(rel_index.PRODUCT, 14, 14, 1, 1.9, 3.8, 15.2)
    assert nb == 50
    np.testing.assert_array_equal(ids, m.bucket_ids(N))

    def oracle(device, autocast):
        P = {k: v.detach().to(device).clone().requires_grad_(True) for k, v in m.named_parameters()}
        xr = x.to(device).clone().requires_grad_(True)
        tab = lambda w: P.get(f"rpe_{w}.lookup_table_weight")
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = vo.rpe_attention(xr, P["qkv.weight"], P["qkv.bias"], P["proj.weight"], P["proj.bias"], heads, ids,
                                 rpe_q=tab("q"), rpe_k=tab("k"), rpe_v=tab("v"), mode="contextual")
        y.float().backward(gy.to(device))
        return y.detach().float().cpu(), xr.grad.float().cpu(), {k: v.grad.float().cpu() for k, v in P.items()}

    y_ref, gx_ref, gp_ref = oracle("cpu", False)
    y_cmp, gx_cmp, gp_cmp = oracle("cuda", True)
    xc = x.cuda().requires_grad_(True)
    y = m(xc)
    y.backward(gy.cuda().to(y.dtype))
    pairs = [("y", y.detach().float().cpu(), y_cmp, y_ref), ("dx", xc.grad.float().cpu(), gx_cmp, gx_ref)]
    pairs += [(pn, p.grad.float().cpu(), gp_cmp[pn], gp_ref[pn]) for pn, p in m.named_parameters()]
    for name, ours, cmp_, ref in pairs:
        a, c = rel_err(ours, ref), rel_err(cmp_, ref)
        print(f"[c2 block rpe_on={rpe_on}] {name}: cream {a:.3e}  torch autocast(bf16) {c:.3e}")
        assert a <= max(RATIO * c, 2.0 ** -8), f"{name}: {a:.3e} vs comparator {c:.3e}"


def test_attention_large_magnitude_tables():
    """The kernel stages R = Q . T^T in fp16 (11-bit mantissa, range 65504): tables of magnitude ~10
    and queries of magnitude ~2 give |R| of a few hundred - far inside the range - and must keep the
    forward / backward parity of the small-table cases."""
    from cream_b200.autoformer.functional import AutoformerAttentionFn
    from cream_b200 import ops
    B, N, h = 2, 197, 3
    bf = lambda t: t.to(torch.bfloat16).float()
    qkv = ops.empty_bf16(B * N, 3 * 64 * h)
    qkv.copy_(rand((B * N, 3 * 64 * h), 61, 2.0))
    qkv = qkv.reshape(B, N, -1).requires_grad_(True)
    tabs = [bf(rand((30, 64), 70 + i, 10.0 if i < 2 else 1.0)).cuda().requires_grad_(True) for i in range(4)]
    out = AutoformerAttentionFn.apply(qkv, h, 0.125, 14, *tabs)
    dout = bf(rand((B, N, 64 * h), 62)).cuda()
    out.backward(dout.to(out.dtype))
    q_ref = qkv.detach().float().cpu().requires_grad_(True)
    t_ref = [t.detach().cpu().requires_grad_(True) for t in tabs]
    ref = vo.attention_core_autoformer(q_ref.reshape(B, N, 3, h, 64), tuple(t_ref), 14, 0.125)
    ref.backward(dout.cpu())
    assert torch.isfinite(out).all()
    # logits of magnitude ~100 make the softmax nearly one-hot: the output is dominated by a few
    # value rows and errors are those of the bf16 probabilities
    assert rel_err(out.float().cpu(), ref.detach()) < 8e-3
    assert rel_err(qkv.grad.float().cpu(), q_ref.grad) < 3e-2


# ------------------------------------------------------------------------------------------------
# (iv) SupernetTrainer.step x 3 against the oracle's AdamW loop (the loop bench.cpu_baseline times)
# ------------------------------------------------------------------------------------------------
def test_trainer_three_steps_follow_the_oracle_adamw_loop():
    from cream_b200 import ops
    from cream_b200.trainer import SupernetTrainer
    ops.SHADOWS.clear()
    spec, space = vo.SUPERNET_T, vo.SEARCH_SPACE["T"]
    batch, steps, lr, wd = 8, 3, 1e-3, 0.05
    sd = vo.init_params(spec, seed=3)
    images = [rand((batch, 3, 224, 224), seed=40 + s) for s in range(steps)]
    targets = [torch.from_numpy(np.random.default_rng(50 + s).integers(0, 1000, batch)) for s in range(steps)]

    def adamw_groups(named):
        decay = [p for n, p in named.items() if not (p.ndim <= 1 or n.endswith(".bias") or n in ('pos_embed', 'cls_token'))]
        rest = [p for n, p in named.items() if (p.ndim <= 1 or n.endswith(".bias") or n in ('pos_embed', 'cls_token'))]
        return [{"params": decay, "weight_decay": wd}, {"params": rest, "weight_decay": 0.0}]

    def torch_loop(device, autocast):
        P = {k: v.to(device).clone().requires_grad_(True) for k, v in sd.items()}
        opt = torch.optim.AdamW(adamw_groups(P), lr=lr)
        rnd, losses = random.Random(0), []
        for s in range(steps):
            cfg = vo.sample_configs(space, rnd)
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                logits = vo.supernet_forward(P, cfg, images[s].to(device), spec)
            loss = F.cross_entropy(logits.float(), targets[s].to(device))
            loss.backward()
            opt.step()
            losses.append(float(loss))
        return losses, {k: v.detach().float().cpu() for k, v in P.items()}

    ref_losses, ref_P = torch_loop("cpu", False)
    cmp_losses, cmp_P = torch_loop("cuda", True)
    net = _mirror(spec)
    net.load_state_dict(sd)
    net = net.cuda().train()
    tr = SupernetTrainer(net, space, lr=lr, weight_decay=wd)
    rnd, losses = random.Random(0), []
    for s in range(steps):
        losses.append(float(tr.step(images[s].cuda(), targets[s].cuda(), rnd=rnd)))
    print(f"\n[trainer] losses  oracle {ref_losses}\n[trainer]         cream  {losses}\n[trainer]    autocast  {cmp_losses}")
    for a, b, c in zip(losses, ref_losses, cmp_losses):
        assert abs(a - b) <= max(RATIO * abs(c - b), 5e-3), "loss trajectory"
    # parameter updates: Adam's normalised step is +-lr wherever a gradient exists, so noise-level
    # gradient elements may flip sign under ANY bf16 path; compare the update DIRECTION with the fp32
    # loop's and hold it to the stock mixed-precision path's own agreement
    num = den_a = den_b = num_c = den_c = 0.0
    for k, p in net.named_parameters():
        d_ref = (ref_P[k] - sd[k]).double().flatten()
        d_our = (p.detach().float().cpu() - sd[k]).double().flatten()
        d_cmp = (cmp_P[k] - sd[k]).double().flatten()
        num += float(d_our @ d_ref); den_a += float(d_our @ d_our); den_b += float(d_ref @ d_ref)
        num_c += float(d_cmp @ d_ref); den_c += float(d_cmp @ d_cmp)
        # identity layers (grad None) and un-sampled slices of un-decayed parameters: exactly unchanged in BOTH
        # torch loops (an isolated exact zero of the fp32 loop, e.g. a key-bias gradient that cancels
        # exactly, is noise-level and moves under any reduced-precision path)
        untouched = (d_ref == 0) & (d_cmp == 0)
        if untouched.any():
            bad = d_our[untouched].abs()
            assert float(bad.max()) == 0.0, (f"{k}: update outside the sampled slices: {int((bad > 0).sum())} of "
                                             f"{int(untouched.sum())} untouched elements moved, max {float(bad.max()):.3e}")
    cos_our, cos_cmp = num / (den_a * den_b) ** 0.5, num_c / (den_c * den_b) ** 0.5
    print(f"[trainer] cosine(update, fp32 update): cream {cos_our:.4f}   autocast(bf16) {cos_cmp:.4f}")
    assert cos_our >= cos_cmp - 0.02


# ------------------------------------------------------------------------------------------------
# (v) 2-rank NCCL gradients == 1-rank double-batch gradients
# ------------------------------------------------------------------------------------------------
def test_two_rank_nccl_gradients_equal_one_rank_double_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2; scripts/ddp_grad_check.py is the torchrun script)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(ROOT / "scripts" / "ddp_grad_check.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DDP_GRAD_CHECK_OK" in r.stdout
