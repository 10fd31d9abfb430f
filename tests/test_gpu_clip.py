"""TinyCLIP towers and contrastive step on the GPU against the reference's own model.py / loss.py
(loaded unmodified through oracle/refload.py from $CREAM_REFERENCE, /root/reference or baseline/_ref).

Comparator: the reference in fp32 is the truth; the reference under torch.autocast(bf16) - the
arithmetic `--precision amp` trains with - gives the error a correct bf16 implementation is allowed.
cream_b200 must be within 1.5x of that error (plus a small floor) on features and on every gradient.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import refload
from tests.helpers import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refload.available(), reason="reference not staged")]

SMALL = dict(embed_dim=64, vision_cfg=dict(image_size=64, layers=2, width=128, patch_size=32),
             text_cfg=dict(context_length=16, vocab_size=100, width=128, heads=2, layers=2))
MID = dict(embed_dim=128, vision_cfg=dict(image_size=224, layers=3, width=256, patch_size=32),
           text_cfg=dict(context_length=77, vocab_size=1000, width=192, heads=3, layers=3))


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cream_b200 import _lib, ops
    _lib.load()
    ops.SHADOWS.clear()
    yield
    torch.cuda.synchronize()


def _ref_clip(cfg, seed):
    torch.manual_seed(seed)
    m = refload.open_clip_model()
    return m.CLIP(cfg["embed_dim"], dict(cfg["vision_cfg"]), dict(cfg["text_cfg"])).cuda()


def _batch(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    s = cfg["vision_cfg"]["image_size"]
    L, V = cfg["text_cfg"]["context_length"], cfg["text_cfg"]["vocab_size"]
    images = torch.randn(B, 3, s, s, generator=g)
    text = torch.randint(1, V - 1, (B, L), generator=g)
    eot = torch.randint(2, L, (B,), generator=g)
    for b in range(B):
        text[b, eot[b]] = V - 1           # the eot token is the largest id (model.py:795-797)
        text[b, eot[b] + 1:] = 0
    return images.cuda(), text.cuda()


def _loss_and_grads(model, loss_fn, images, text, autocast=False):
    model.zero_grad(set_to_none=True)
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            fi, ft, s = model(images, text)
    else:
        fi, ft, s = model(images, text)
    loss = loss_fn(fi.float(), ft.float(), s.float())
    loss.backward()
    grads = {k: p.grad.detach().float().clone() for k, p in model.named_parameters() if p.grad is not None}
    return fi.detach().float(), ft.detach().float(), loss.detach().float(), grads


@pytest.mark.parametrize("cfg,B", [(SMALL, 6), (MID, 16)], ids=["small", "mid"])
def test_towers_and_loss_against_the_reference_model(cfg, B):
    from cream_b200 import clip
    ref = _ref_clip(cfg, seed=11)
    with torch.no_grad():                  # non-trivial gains / biases so their gradients carry signal
        for n, p in ref.named_parameters():
            if p.ndim < 2 and "logit_scale" not in n:
                p.add_(0.1 * torch.randn_like(p))
    ours = clip.CLIP(cfg["embed_dim"], cfg["vision_cfg"], cfg["text_cfg"]).cuda()
    ours.load_state_dict(ref.state_dict(), strict=True)
    images, text = _batch(cfg, B, seed=5)
    ref_loss = refload.open_clip_loss().ClipLoss()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        fi0, ft0, l0, g0 = _loss_and_grads(ref, ref_loss, images, text)
        fi1, ft1, l1, g1 = _loss_and_grads(ref, ref_loss, images, text, autocast=True)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    fi2, ft2, l2, g2 = _loss_and_grads(ours, clip.ClipLoss(), images, text)

    def within(name, got, amp, truth, floor):
        e_amp, e_got = rel_err(amp, truth), rel_err(got, truth)
        assert e_got <= 1.5 * e_amp + floor, f"{name}: cream {e_got:.3e} vs autocast reference {e_amp:.3e}"
        return e_got, e_amp

    within("image features", fi2, fi1, fi0, 1e-3)
    within("text features", ft2, ft1, ft0, 1e-3)
    assert abs(float(l2) - float(l0)) <= 1.5 * abs(float(l1) - float(l0)) + 2e-3
    assert set(g2) == set(g0)
    worst = 0.0
    for k in sorted(g0):
        if float(g0[k].norm()) < 1e-7:
            continue
        e, _ = within("grad " + k, g2[k], g1[k], g0[k], 5e-3)
        worst = max(worst, e)
    print(f"[clip {cfg['vision_cfg']['width']}/{cfg['text_cfg']['width']}] worst gradient rel err {worst:.3e}")


def test_text_tower_respects_the_causal_mask_and_eot_pooling():
    """Tokens after the eot position must not influence the feature (causal mask, model.py:756-762 +
    eot pooling, model.py:795-797)."""
    from cream_b200 import clip
    torch.manual_seed(2)
    m = clip.CLIP(SMALL["embed_dim"], SMALL["vision_cfg"], SMALL["text_cfg"]).cuda()
    _, text = _batch(SMALL, 4, seed=9)
    a = m.encode_text(text)
    eot = text.argmax(-1)
    text2 = text.clone()
    for b in range(4):
        text2[b, eot[b] + 1:] = torch.randint(1, 50, (text.shape[1] - int(eot[b]) - 1,), device="cuda")
    b_ = m.encode_text(text2)
    assert torch.equal(a, b_)


def test_vit_b_32_step_runs_and_matches_reference_features():
    """BASELINE config 4's model (ViT-B/32 image tower 12 x 768, text tower 12 x 512, ctx 77)."""
    from cream_b200 import clip
    cfg = clip.VIT_B_32
    ref = _ref_clip(cfg, seed=1)
    ours = clip.CLIP(cfg["embed_dim"], cfg["vision_cfg"], cfg["text_cfg"]).cuda()
    ours.load_state_dict(ref.state_dict(), strict=True)
    images, text = _batch(cfg, 8, seed=3)
    with torch.no_grad():
        fi0, ft0, _ = ref(images, text)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            fi1, ft1, _ = ref(images, text)
    fi2, ft2, s = ours(images, text)
    assert rel_err(fi2, fi0) <= 1.5 * rel_err(fi1.float(), fi0) + 1e-3
    assert rel_err(ft2, ft0) <= 1.5 * rel_err(ft1.float(), ft0) + 1e-3
    loss = clip.ClipLoss()(fi2, ft2, s)
    loss.backward()
    for n, p in ours.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n


def test_contrastive_trainer_follows_the_reference_trajectory():
    """3 optimiser steps: ClipTrainer vs the reference model + reference loss + torch AdamW with the
    reference's parameter groups (training/optimizer.py:22-52)."""
    from cream_b200 import clip
    cfg = SMALL
    ref = _ref_clip(cfg, seed=4)
    ours = clip.CLIP(cfg["embed_dim"], cfg["vision_cfg"], cfg["text_cfg"]).cuda()
    ours.load_state_dict(ref.state_dict(), strict=True)
    init = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    tr = clip.ClipTrainer(ours, lr=1e-3)
    named = list(ref.named_parameters())
    skip = lambda n, p: p.ndim < 2 or "bn" in n or "ln" in n or "bias" in n or "logit_scale" in n
    opt = torch.optim.AdamW([dict(params=[p for n, p in named if skip(n, p)], weight_decay=0.0),
                             dict(params=[p for n, p in named if not skip(n, p)], weight_decay=0.2)],
                            lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    ref_loss = refload.open_clip_loss().ClipLoss()
    for step in range(3):
        images, text = _batch(cfg, 8, seed=20 + step)
        opt.zero_grad(set_to_none=True)
        fi, ft, s = ref(images, text)
        want = ref_loss(fi, ft, s)
        want.backward()
        opt.step()
        with torch.no_grad():
            ref._logit_scale.logit_scale.clamp_(0, math.log(100))
        got = tr.step(images, text)
        assert abs(float(got) - float(want)) <= 2e-2 * max(1.0, abs(float(want))), (step, float(got), float(want))
    sd = ref.state_dict()
    for k, v in ours.state_dict().items():
        if v.ndim < 2:
            continue
        moved = float((sd[k] - init[k]).norm())
        assert moved > 0, k
        # Adam normalises every element's step, so elements whose gradient is bf16 noise may step the
        # other way; the update as a whole must still agree with the fp32 reference's
        assert float((v - sd[k]).norm()) <= 0.5 * moved, (k, float((v - sd[k]).norm()), moved)


def test_graph_replay_equals_the_eager_step():
    """ClipTrainer(graph=True): the captured step replayed three times against the same three steps run eagerly from
    the same initial weights - same losses and parameters up to the run-to-run re-association of the split-K reduces -
    and the warm-up before the capture must not advance the training state."""
    from cream_b200 import clip
    cfg = SMALL
    torch.manual_seed(0)
    a = clip.CLIP(cfg["embed_dim"], cfg["vision_cfg"], cfg["text_cfg"]).cuda()
    b = clip.CLIP(cfg["embed_dim"], cfg["vision_cfg"], cfg["text_cfg"]).cuda()
    b.load_state_dict(a.state_dict())
    ta, tb = clip.ClipTrainer(a, lr=1e-3, graph=True), clip.ClipTrainer(b, lr=1e-3)
    for step in range(3):
        images, text = _batch(cfg, 8, seed=40 + step)
        la, lb = ta.step(images, text), tb.step(images, text)
        assert abs(float(la) - float(lb)) <= 2e-3 * max(1.0, abs(float(lb))), (step, float(la), float(lb))
    assert ta._graph is not None, "the capture was refused"
    for (n, p), q in zip(a.named_parameters(), b.parameters()):
        assert rel_err(p, q) < 2e-3, n
