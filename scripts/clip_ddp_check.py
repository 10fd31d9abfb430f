#!/usr/bin/env python
"""torchrun --nproc-per-node 2 scripts/clip_ddp_check.py

TinyCLIP contrastive step over N ranks (NCCL): with the features gathered across ranks (local loss +
gather with gradient, and the global-loss variant) the rank-averaged gradients must equal the gradients ONE rank
computes on the concatenated batch, and the replicas must stay identical through optimizer steps.  Every rank builds
the model under a different seed (covers the construction-time broadcast).  Prints CLIP_DDP_CHECK_OK on rank 0."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CFG = dict(embed_dim=128, vision_cfg=dict(image_size=224, layers=3, width=256, patch_size=32),
           text_cfg=dict(context_length=77, vocab_size=1000, width=192, heads=3, layers=3))


def batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(n, 3, 224, 224, generator=g)
    text = torch.randint(1, 998, (n, 77), generator=g)
    eot = torch.randint(2, 77, (n,), generator=g)
    for b in range(n):
        text[b, eot[b]] = 999
        text[b, eot[b] + 1:] = 0
    return images, text


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from cream_b200 import clip

    torch.manual_seed(500 + rank)
    net = clip.CLIP(CFG["embed_dim"], CFG["vision_cfg"], CFG["text_cfg"]).to(dev).train()
    with torch.no_grad():
        for n, p in net.named_parameters():
            if p.ndim < 2 and "logit_scale" not in n:
                p.add_(0.05 * torch.randn_like(p))
    per = 8
    images, text = batch(per * world, 11)
    worst = 0.0
    for local_loss, with_grad in ((True, True), (False, True)):
        tr = clip.ClipTrainer(net, lr=1e-3, local_loss=local_loss, gather_with_grad=with_grad)   # broadcasts rank 0's weights
        sl = slice(rank * per, (rank + 1) * per)
        net.zero_grad(set_to_none=True)
        fi, ft, s = net(images[sl].to(dev), text[sl].to(dev))
        loss = tr.loss(fi, ft, s)
        loss.backward()
        grads = {}
        for n, p in net.named_parameters():
            g = p.grad.detach().clone()
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
            grads[n] = g
        # one rank, whole batch, no collectives
        net.zero_grad(set_to_none=True)
        fi, ft, s = net(images.to(dev), text.to(dev))
        full = clip.ClipLoss()(fi, ft, s)
        full.backward()
        tot = loss.detach().clone()
        dist.all_reduce(tot, op=dist.ReduceOp.AVG)
        assert abs(float(tot) - float(full)) <= 2e-3 * abs(float(full)), (local_loss, float(tot), float(full))
        for n, p in net.named_parameters():
            ref = p.grad.detach()
            scale = float(ref.norm())
            if scale < 1e-8:
                continue
            e = float((grads[n] - ref).norm()) / scale
            worst = max(worst, e)
            # bf16 operands: the two evaluations differ in GEMM tiling (M = 8 vs 16 rows per rank) only through
            # accumulation order; features are rounded identically
            assert e <= 2e-2, (local_loss, with_grad, n, e)
    # two optimizer steps; replicas must stay identical
    tr = clip.ClipTrainer(net, lr=1e-3)
    for step in range(2):
        im, tx = batch(per * world, 30 + step)
        tr.step(im[rank * per:(rank + 1) * per].to(dev), tx[rank * per:(rank + 1) * per].to(dev))
    digest = torch.stack([p.detach().double().sum() for p in net.parameters()])
    both = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(both, digest)
    assert all(torch.equal(b, both[0]) for b in both), "replicas diverged"
    if rank == 0:
        print(f"CLIP_DDP_CHECK_OK world={world} worst_rel_err={worst:.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
