"""Repeat tests/test_gpu_native.py::test_staged_host_batches_give_the_same_steps_as_resident_ones and print the loss
trajectories (run-to-run spread of the resident arm vs the staged arm).  CREAM_PDL=0/1 in the environment."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from cream_b200.deit import DeitIrpe, DeitTrainer
g = torch.Generator().manual_seed(3)
host = [(torch.randn(4, 3, 224, 224, generator=g).pin_memory(), torch.randint(0, 1000, (4,), generator=g).pin_memory()) for _ in range(3)]
def run(staged):
    torch.manual_seed(0)
    net = DeitIrpe(depth=2, drop_path_rate=0.0).cuda().train()
    tr = DeitTrainer(net, lr=1e-3)
    out = []
    nxt = tr.stage(*host[0]) if staged else None
    for s in range(5):
        x, y = host[s % 3]
        if staged:
            cur = nxt
            loss = tr.step(cur)
            if s + 1 < 5:
                nxt = tr.stage(*host[(s + 1) % 3])
        else:
            loss = tr.step(x.cuda(), y.cuda())
        out.append(float(loss))
    return np.array(out)
base = run(False)
print("PDL", os.environ.get("CREAM_PDL", "1"), "resident[0]", base)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    r, s = run(False), run(True)
    print(f"  rep {i}: resident-vs-first {np.max(np.abs(r - base) / np.abs(base)):.2e}   staged-vs-first {np.max(np.abs(s - base) / np.abs(base)):.2e}  staged {s}")
