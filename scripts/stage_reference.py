#!/usr/bin/env python
"""Stage the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored, NOT
gpurun-ignored) so that the GPU box — which has no /root/reference — can run the reference's own
supernet_transformer.py / supernet_engine.py / irpe.py / rpe_vision_transformer.py over the
cream_b200 drop-ins, and time the reference path on the same B200.

    python scripts/stage_reference.py            # copies from $CREAM_REFERENCE or /root/reference

Nothing staged here is product source or enters git history; tests skip when it is absent.  The
files are byte-for-byte copies (checked by tests/test_reference_staged.py through sha256 of the
originals when the checkout is present)."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("CREAM_REFERENCE", "/root/reference"))
DST = ROOT / "baseline" / "_ref"

FILES = [
    "AutoFormer/model/supernet_transformer.py", "AutoFormer/model/utils.py",
    "AutoFormer/model/module/Linear_super.py", "AutoFormer/model/module/embedding_super.py",
    "AutoFormer/model/module/layernorm_super.py", "AutoFormer/model/module/multihead_super.py",
    "AutoFormer/model/module/qkv_super.py",
    "AutoFormer/supernet_engine.py", "AutoFormer/lib/utils.py",
    "AutoFormer/experiments/supernet/supernet-T.yaml", "AutoFormer/experiments/supernet/supernet-S.yaml",
    "AutoFormer/experiments/supernet/supernet-B.yaml",
    "iRPE/DeiT-with-iRPE/irpe.py", "iRPE/DeiT-with-iRPE/rpe_vision_transformer.py",
    "iRPE/DeiT-with-iRPE/rpe_models.py", "iRPE/DeiT-with-iRPE/models.py",
    "iRPE/DeiT-with-iRPE/rpe_ops/rpe_index.py",
    "iRPE/DETR-with-iRPE/models/rpe_attention/irpe.py",
    "iRPE/DETR-with-iRPE/models/rpe_attention/rpe_attention_function.py",
    "iRPE/DETR-with-iRPE/models/rpe_attention/multi_head_attention.py",
    "iRPE/DETR-with-iRPE/models/rpe_attention/__init__.py",
    "TinyViT/models/tiny_vit.py",
    "TinyCLIP/src/open_clip/model.py", "TinyCLIP/src/open_clip/loss.py",
]


def main() -> int:
    if not REF.exists():
        print(f"[stage_reference] {REF} not present; nothing staged")
        return 0
    manifest = {}
    for rel in FILES:
        src = REF / rel
        if not src.exists():
            print(f"[stage_reference] missing in the reference: {rel}")
            continue
        dst = DST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(src.read_bytes()).hexdigest()
    (DST / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True))
    print(f"[stage_reference] {len(manifest)} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
