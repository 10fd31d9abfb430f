"""Shared test helpers: golden-fixture comparison and seeded synthetic inputs."""
from __future__ import annotations

import numpy as np
import torch


def rand(shape, seed, scale=1.0):
    """Same generator as tests/golden/make_golden.py (numpy PCG64, platform independent)."""
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32) * scale)


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().double().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


def rel_err(a, b) -> float:
    a = _np(a)
    b = _np(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def check_summary(golden, prefix: str, value, rtol: float, what: str = ""):
    """Compare `value` with a fixture written by make_golden.summarize (full tensor, or
    sum / abs-sum / strided sample for large ones).  rtol is a relative L2 tolerance."""
    a = value.detach().double().cpu().numpy() if isinstance(value, torch.Tensor) else np.asarray(value, np.float64)
    if prefix + "_full" in golden:
        ref = golden[prefix + "_full"]
        assert a.shape == ref.shape, f"{what or prefix}: shape {a.shape} vs {ref.shape}"
        e = rel_err(a, ref)
        assert e <= rtol, f"{what or prefix}: rel L2 err {e:.3e} > {rtol:.1e}"
        return e
    flat = a.reshape(-1)
    ref_s = golden[prefix + "_sample"]
    e = rel_err(flat[::101], ref_s)
    assert e <= rtol, f"{what or prefix}: sample rel L2 err {e:.3e} > {rtol:.1e}"
    abssum = float(golden[prefix + "_abssum"])
    e2 = abs(np.abs(flat).sum() - abssum) / max(abssum, 1e-30)
    assert e2 <= rtol, f"{what or prefix}: abs-sum rel err {e2:.3e} > {rtol:.1e}"
    return max(e, e2)
