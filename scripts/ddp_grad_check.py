#!/usr/bin/env python
"""torchrun --nproc-per-node 2 scripts/ddp_grad_check.py

N-rank NCCL gradients of SupernetTrainer (per-layer buckets, AVG all-reduce overlapped with backward)
must equal the gradients one rank computes on the concatenated batch (what DDP guarantees,
supernet_train.py:288).  Every rank builds the model under a DIFFERENT seed (the reference seeds with
args.seed + rank), so the check also covers the construction-time parameter broadcast.  Also runs two
optimizer steps and checks that the replicas stay bit-identical.  Prints DDP_GRAD_CHECK_OK on rank 0."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from cream_b200.autoformer.model.supernet_transformer import Vision_TransformerSuper
    from cream_b200.configs import SEARCH_SPACE, SUPERNETS
    from cream_b200.trainer import SupernetTrainer

    spec = SUPERNETS["T"]
    torch.manual_seed(1000 + rank)
    net = Vision_TransformerSuper(img_size=224, patch_size=16, embed_dim=spec["embed_dim"], depth=spec["depth"],
                                  num_heads=spec["num_heads"], mlp_ratio=spec["mlp_ratio"], qkv_bias=True,
                                  drop_path_rate=0.0, gp=True, relative_position=True, change_qkv=True,
                                  max_relative_position=14).to(dev).train()
    with torch.no_grad():   # non-trivial biases / norms so that every gradient is exercised
        for n, p in net.named_parameters():
            if p.ndim <= 1:
                p.add_(0.05 * torch.randn_like(p))
    tr = SupernetTrainer(net, SEARCH_SPACE["T"], lr=1e-3)
    g = torch.Generator().manual_seed(7)
    per = 4
    images = torch.randn(per * world, 3, 224, 224, generator=g)
    targets = torch.randint(0, 1000, (per * world,), generator=g)
    cfg = {"layer_num": 13, "embed_dim": [216] * 13, "num_heads": [3, 4] * 6 + [3], "mlp_ratio": [3.5, 4.0] * 6 + [4.0]}
    tr.forward_backward(images[rank * per:(rank + 1) * per].to(dev), targets[rank * per:(rank + 1) * per].to(dev), config=cfg)
    tr.sync_grads_to_params()
    torch.cuda.synchronize()
    ddp = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in net.named_parameters()}
    # one rank, the whole batch, plain autograd over the same engine (no buckets, no collectives)
    for p in net.parameters():
        p.grad = None
    net.set_sample_config(cfg)
    loss = F.cross_entropy(net(images.to(dev)).float(), targets.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    worst = 0.0
    for n, p in net.named_parameters():
        if p.grad is None:
            assert ddp[n] is None, f"{n}: identity-layer parameter received a gradient under DDP"
            continue
        assert ddp[n] is not None, n
        err = float((ddp[n].double() - p.grad.double()).norm() / p.grad.double().norm().clamp_min(1e-30))
        worst = max(worst, err)
        assert err < 2e-4, f"{n}: {world}-rank gradient differs from the full-batch gradient: {err:.3e}"
    # replicas stay identical through optimizer steps
    for s in range(2):
        tr.step(images[rank * per:(rank + 1) * per].to(dev), targets[rank * per:(rank + 1) * per].to(dev), config=cfg)
    digest = torch.stack([p.detach().double().sum() for p in net.parameters()])
    both = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(both, digest)
    assert all(torch.equal(b, both[0]) for b in both), "replicas diverged"
    if rank == 0:
        print(f"DDP_GRAD_CHECK_OK world={world} worst_rel_err={worst:.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
