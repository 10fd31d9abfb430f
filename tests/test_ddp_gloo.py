"""Host-side logic of the data-parallel path on CPU (gloo, world_size 2): identical subnet
sampling on every rank, per-layer flat gradient buckets, averaged all-reduce, and the
"un-sampled layers keep grad None" contract (supernet_train.py:288, supernet_engine.py:36)."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


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
        from cream_b200 import engine
        from cream_b200.autoformer.model.supernet_transformer import Vision_TransformerSuper
        from cream_b200.trainer import SupernetTrainer, sample_configs
        torch.manual_seed(100 + rank)      # the reference seeds with args.seed + rank (supernet_train.py:197-198)
        model = Vision_TransformerSuper(img_size=64, embed_dim=128, depth=3, num_heads=2, mlp_ratio=4.0,
                                        qkv_bias=True, gp=True, relative_position=True, change_qkv=True)
        choices = dict(mlp_ratio=[3.0, 3.5, 4.0], num_heads=[1, 2], depth=[2, 3], embed_dim=[64, 96, 128])
        tr = SupernetTrainer(model, choices)
        assert tr.world == world
        # 0. replicas start from rank 0's weights whatever the per-rank seed was (DDP's init broadcast)
        digest = torch.stack([p.detach().double().sum() for p in model.parameters()])
        both = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(both, digest)
        assert all(torch.equal(b, both[0]) for b in both), "parameters differ across ranks after construction"
        # 1. the config stream is identical on every rank (the engine seeds `random` with the epoch)
        random.seed(3)
        cfgs = [sample_configs(choices) for _ in range(5)]
        gathered = [None] * world
        dist.all_gather_object(gathered, cfgs)
        assert all(g == gathered[0] for g in gathered)
        # 2. bucket views alias the flat buffers and cover every parameter exactly once
        names = set(dict(model.named_parameters()))
        assert set(tr.buckets.views) == names
        assert sum(len(v) for v in tr.buckets.groups.values()) == len(names)
        for n, p in model.named_parameters():
            assert tr.buckets.views[n].shape == p.shape
        # 3. averaged all-reduce of each group
        cfg = {"layer_num": 2, "embed_dim": [64, 64], "num_heads": [1, 1], "mlp_ratio": [3.5, 4.0]}
        used = ["embed", "head", "block0", "block1"]
        for g in used:
            tr.buckets.flat[g].fill_(float(rank + 1))
            tr._allreduce(g)
        want = sum(range(1, world + 1)) / world
        for g in used:
            assert torch.allclose(tr.buckets.flat[g], torch.full_like(tr.buckets.flat[g], want))
        assert float(tr.buckets.flat["block2"].abs().sum()) == 0.0      # un-sampled layer untouched
        # 4. grads are assigned only to sampled parameters
        sampled = set(engine.sampled_param_names(model.engine_geometry(), cfg))
        for n, p in tr.params.items():
            p.grad = tr.buckets.views[n] if n in sampled else None
        assert model.blocks[2].fc1.weight.grad is None and model.blocks[0].fc1.weight.grad is not None
        tr.optimizer.step()                                              # None grads are skipped
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_gradient_sync_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
