"""GPU debug: DeiT + iRPE native runtime, gradient of the iRPE table vs the reference VisionTransformer."""
import sys
from functools import partial
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from oracle import refload
from tests.helpers import rand, rel_err
from cream_b200.deit import DeitIrpe
vit = refload.rpe_vision_transformer("reference")
cfg = vit.irpe.get_rpe_config(ratio=1.9, method='product', mode='ctx', shared_head=True, skip=1, rpe_on='k')
depth, B = 2, 4
ref = vit.VisionTransformer(patch_size=16, embed_dim=384, depth=depth, num_heads=6, mlp_ratio=4, qkv_bias=True,
                            norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), rpe_config=cfg)
ours = DeitIrpe(depth=depth, drop_path_rate=0.0)
seed = 900
with torch.no_grad():
    for n, p in ours.named_parameters():
        seed += 1
        if n.endswith("norm1.weight") or n.endswith("norm2.weight") or n == "norm.weight":
            p.copy_(1.0 + rand(tuple(p.shape), seed, 0.1))
        else:
            p.copy_(rand(tuple(p.shape), seed, 0.05 if n.endswith(".bias") else 0.02))
ref.load_state_dict(ours.state_dict())
images = rand((B, 3, 224, 224), seed=901)
targets = torch.from_numpy(np.random.default_rng(902).integers(0, 1000, B))
ref.train(); F.cross_entropy(ref(images), targets).backward()
ours = ours.cuda().train()
F.cross_entropy(ours(images.cuda()), targets.cuda()).backward()
R = dict(ref.named_parameters())
for n, p in ours.named_parameters():
    e = rel_err(p.grad.cpu(), R[n].grad)
    flag = "  <<<<" if e > 4e-2 else ""
    print(f"{n:45s} {e:.3e}{flag}")
for i in range(depth):
    n = f"blocks.{i}.attn.rpe_k.lookup_table_weight"
    a, b = dict(ours.named_parameters())[n].grad.cpu()[0], R[n].grad[0]      # (64, 50)
    per = (a - b).norm(dim=0) / b.norm(dim=0).clamp_min(1e-12)
    print(n, "per-bucket rel err:", [round(float(v), 3) for v in per])
    print("   ours/ref norm ratio per bucket:", [round(float(v), 3) for v in a.norm(dim=0) / b.norm(dim=0).clamp_min(1e-12)])
