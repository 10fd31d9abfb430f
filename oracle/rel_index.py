"""ORACLE (test infrastructure only — never imported by the product path).

numpy restatement of the integer index tables on the relative-position attention path.
Every function cites the reference lines it follows (paths relative to the reference
checkout).  Pinned against the reference itself: tests/golden/make_golden.py imports the
reference modules in the build container and commits their outputs as fixtures;
tests/test_oracle_golden.py checks this file against those fixtures.
"""
from __future__ import annotations

import math

import numpy as np

# method ids, iRPE/DeiT-with-iRPE/irpe.py:117-127
EUCLIDEAN, QUANT, PRODUCT, CROSS_ROWS, CROSS_COLS = 0, 1, 3, 41, 42


def autoformer_rel_index(grid: int, max_rel: int = 14):
    """AutoFormer/model/module/multihead_super.py:40-59 (RelativePosition2D_super.forward).

    Returns (idx_v, idx_h): (n, n) int64 with n = grid*grid + 1; row/col 0 (cls) are 0,
    patch pairs are clamp(delta, +-max_rel) + max_rel + 1.
    """
    length = grid * grid
    g = int(length ** 0.5)
    rq = np.arange(length)
    rk = np.arange(length)
    dv = rk[None, :] // g - rq[:, None] // g
    dh = rk[None, :] % g - rq[:, None] % g
    dv = np.clip(dv, -max_rel, max_rel) + max_rel + 1
    dh = np.clip(dh, -max_rel, max_rel) + max_rel + 1
    dv = np.pad(dv, ((1, 0), (1, 0)), constant_values=0)
    dh = np.pad(dh, ((1, 0), (1, 0)), constant_values=0)
    return dv.astype(np.int64), dh.astype(np.int64)


def piecewise_index(rel: np.ndarray, alpha: float, beta: float, gamma: float) -> np.ndarray:
    """iRPE/DeiT-with-iRPE/irpe.py:19-52.

    Integer input: identity inside |x| <= alpha; outside, the log branch is evaluated in
    float32 exactly as torch does for an int64 tensor divided by a Python float, then
    round-half-even, clip(max=beta) and a truncating cast.  Float input (euclidean
    method) additionally rounds inside the alpha band (irpe.py:45-47).
    """
    rel = np.asarray(rel)
    is_float = np.issubdtype(rel.dtype, np.floating)
    x = rel.astype(np.float32)
    ax = np.abs(x)
    mask = ax <= np.float32(alpha)
    ax_far = np.where(mask, np.float32(1.0), ax)  # the near band never uses the log branch
    y = np.float32(alpha) + np.log(ax_far / np.float32(alpha)).astype(np.float32) / np.float32(
        math.log(gamma / alpha)) * np.float32(beta - alpha)
    y = np.minimum(np.rint(y.astype(np.float32)), np.float32(beta))
    y = (np.sign(x) * y)
    out_far = np.trunc(y).astype(np.int64)
    near = np.rint(x).astype(np.int64) if is_float else rel.astype(np.int64)
    return np.where(mask, near, out_far)


def irpe_num_buckets(method: int, beta: float) -> int:
    """irpe.py:260-283."""
    beta_int = int(beta)
    return (2 * beta_int + 1) ** 2 if method == PRODUCT else 2 * beta_int + 1


def irpe_bucket_ids(method: int, height: int, width: int, skip: int, alpha: float, beta: float,
                    gamma: float):
    """irpe.py:291-415 (get_bucket_ids_2d_without_skip + get_bucket_ids_2d).

    Returns (ids (skip+L, skip+L) int64, num_buckets including the skip bucket).
    """
    rows = np.repeat(np.arange(height)[:, None], width, 1).reshape(-1)
    cols = np.repeat(np.arange(width)[None, :], height, 0).reshape(-1)
    dr = rows[:, None] - rows[None, :]          # pos1 - pos2, irpe.py:339-342
    dc = cols[:, None] - cols[None, :]
    beta_int = int(beta)
    if method == PRODUCT:                        # irpe.py:176-202
        s = 2 * beta_int + 1
        ids = (piecewise_index(dr, alpha, beta, gamma) + beta_int) * s + (
            piecewise_index(dc, alpha, beta, gamma) + beta_int)
    elif method == EUCLIDEAN:                    # irpe.py:131-149
        dis = np.rint(np.sqrt((dr * dr + dc * dc).astype(np.float32)))
        ids = piecewise_index(dis.astype(np.float32), alpha, beta, gamma) + beta_int
    elif method == QUANT:                        # irpe.py:152-172
        ids = piecewise_index(dr * dr + dc * dc, alpha, beta, gamma) + beta_int
    elif method == CROSS_ROWS:                   # irpe.py:205-224
        ids = piecewise_index(dr, alpha, beta, gamma) + beta_int
    elif method == CROSS_COLS:                   # irpe.py:227-247
        ids = piecewise_index(dc, alpha, beta, gamma) + beta_int
    else:
        raise NotImplementedError(method)
    nb = irpe_num_buckets(method, beta)
    if skip > 0:                                 # irpe.py:401-413
        L = height * width
        full = np.full((skip + L, skip + L), nb, dtype=np.int64)
        full[skip:, skip:] = ids
        ids = full
        nb += 1
    return ids.astype(np.int64), nb


def rpe_index_fwd(inp: np.ndarray, index: np.ndarray) -> np.ndarray:
    """rpe_ops/rpe_index.py:11-39 / rpe_index.cpp:8-73: Y[b,h,i,j] = input[b,h,i,index[i,j]]."""
    lq, lk = index.shape
    return np.take_along_axis(inp, np.broadcast_to(index[None, None], inp.shape[:2] + (lq, lk)), axis=3)


def rpe_index_bwd(grad_out: np.ndarray, index: np.ndarray, num_buckets: int) -> np.ndarray:
    """rpe_index.cpp:82-124: grad_input[b,h,i,index[i,j]] += grad_output[b,h,i,j]."""
    b, h, lq, lk = grad_out.shape
    gi = np.zeros((b, h, lq, num_buckets), dtype=np.float64)
    for i in range(lq):
        np.add.at(gi[:, :, i, :], (slice(None), slice(None), index[i]), grad_out[:, :, i, :])
    return gi.astype(grad_out.dtype)
