"""TinyCLIP towers: host-side surface on CPU.

* parameter names / shapes / buffers equal the reference CLIP's (TinyCLIP/src/open_clip/model.py loaded
  unmodified through oracle/refload.py) so a reference checkpoint loads with strict=True;
* ClipLoss equals the reference ClipLoss (loss.py) on one process;
* world_size-2 gloo: the gathered-feature loss (local_loss / gather_with_grad in every combination the
  reference accepts) reproduces the single-process full-batch loss and gradients;
* the product path refuses CPU tensors.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import refload

SMALL = dict(embed_dim=64, vision_cfg=dict(image_size=64, layers=2, width=128, patch_size=32),
             text_cfg=dict(context_length=16, vocab_size=100, width=128, heads=2, layers=2))

needs_ref = pytest.mark.skipif(not refload.available(), reason="reference not staged")


def _ref_clip(cfg):
    m = refload.open_clip_model()
    return m.CLIP(cfg["embed_dim"], dict(cfg["vision_cfg"]), dict(cfg["text_cfg"]))


@needs_ref
@pytest.mark.parametrize("which", ["small", "ViT-B-32"])
def test_parameter_surface_equals_the_reference(which):
    from cream_b200 import clip
    cfg = SMALL if which == "small" else clip.VIT_B_32
    dev = "cpu" if which == "small" else "meta"
    with torch.device(dev):
        ours = clip.CLIP(cfg["embed_dim"], cfg["vision_cfg"], cfg["text_cfg"])
        ref = _ref_clip(cfg)
    a = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert a == b
    assert {k for k, _ in ours.named_buffers()} == {k for k, _ in ref.named_buffers()}
    if which == "small":
        ours.load_state_dict(ref.state_dict(), strict=True)
        assert torch.equal(ours._text_encoder.attn_mask, ref._text_encoder.attn_mask)
        assert float(ours.logit_scale.detach()) == float(ref._logit_scale.logit_scale.detach())


@needs_ref
def test_clip_loss_equals_the_reference_loss():
    from cream_b200 import clip
    ref_loss = refload.open_clip_loss().ClipLoss()
    g = torch.Generator().manual_seed(3)
    fi = torch.nn.functional.normalize(torch.randn(12, 32, generator=g), dim=-1).requires_grad_()
    ft = torch.nn.functional.normalize(torch.randn(12, 32, generator=g), dim=-1).requires_grad_()
    s = torch.tensor(14.3, requires_grad=True)
    want = ref_loss(fi, ft, s)
    gw = torch.autograd.grad(want, (fi, ft, s))
    got = clip.ClipLoss()(fi, ft, s)
    gg = torch.autograd.grad(got, (fi, ft, s))
    assert torch.equal(want, got)
    for x, y in zip(gw, gg):
        assert torch.equal(x, y)


def test_towers_refuse_cpu_tensors():
    from cream_b200 import clip
    m = clip.CLIP(SMALL["embed_dim"], SMALL["vision_cfg"], SMALL["text_cfg"])
    with pytest.raises(RuntimeError, match="CUDA"):
        m.encode_image(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.encode_text(torch.zeros(1, 16, dtype=torch.long))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cream_b200 import clip
        g = torch.Generator().manual_seed(5)
        B, D = 6, 16
        fi_all = torch.nn.functional.normalize(torch.randn(world * B, D, generator=g, dtype=torch.float64), dim=-1)
        ft_all = torch.nn.functional.normalize(torch.randn(world * B, D, generator=g, dtype=torch.float64), dim=-1)
        scale = torch.tensor(9.0, dtype=torch.float64)
        # single-process statement of the whole batch
        a = fi_all.clone().requires_grad_()
        b = ft_all.clone().requires_grad_()
        full = clip.ClipLoss()(a, b, scale)
        full.backward()
        sl = slice(rank * B, (rank + 1) * B)
        out = {}
        for local_loss in (False, True):
            for with_grad in (False, True):
                fi = fi_all[sl].clone().requires_grad_()
                ft = ft_all[sl].clone().requires_grad_()
                loss = clip.ClipLoss(local_loss=local_loss, gather_with_grad=with_grad, rank=rank, world_size=world)(fi, ft, scale)
                loss.backward()
                # the job's loss is the mean over ranks; local_loss gives each rank its rows' share
                tot = loss.detach().clone()
                dist.all_reduce(tot)
                tot /= world
                assert torch.allclose(tot, full.detach(), atol=1e-12), (local_loss, with_grad, float(tot), float(full))
                if with_grad:
                    # d(mean over ranks of loss_r)/d(features of this rank) == full-batch gradient rows:
                    # all_gather's backward sums the other ranks' contributions, DDP averages over ranks
                    assert torch.allclose(fi.grad / world, a.grad[sl], atol=1e-12), (local_loss, "image grad")
                    assert torch.allclose(ft.grad / world, b.grad[sl], atol=1e-12), (local_loss, "text grad")
                out[(local_loss, with_grad)] = float(loss)
        # the trainer's gradient exchange: one flat all-reduce, averaged, written back with the original shapes
        shapes = [(3, 5), (7,), (2, 2, 2), ()]
        grads = [torch.full(sh, float(rank + 1) * (i + 1), dtype=torch.float64) for i, sh in enumerate(shapes)]
        clip.average_gradients(grads, world)
        mean_rank = sum(range(1, world + 1)) / world
        for i, (gr, sh) in enumerate(zip(grads, shapes)):
            assert gr.shape == torch.Size(sh) and torch.allclose(gr, torch.full(sh, mean_rank * (i + 1), dtype=torch.float64))
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_gathered_contrastive_loss_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
