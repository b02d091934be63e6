"""CPU restatement (numpy, dictionary lookups) of the reference's pooling / refinement ops -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package; the product path never does.

Follows, per function:
  coarsened_ijk      ops/BuildCoarseGridFromFine.cu:50,136     coarse = floor(fine / factor), set semantics
  refined_ijk        ops/BuildFineGridFromCoarse.cu            fine = factor * coarse + [0, factor)^3
  coarse/fine metadata  detail/utils/VoxelSizeUtils.h:13-35
  max_pool / avg_pool   ops/MaxPool.cu:16-63,65-122; ops/AvgPool.cu:17-65,67-110,145  (window = stride * c + [0, factor)^3;
                        max over ACTIVE children; a window with no active child gives 0 (the documented contract,
                        fvdb/nn/modules.py:125-128 -- the kernel's own -inf initialiser, MaxPool.cu:46, is a reference
                        defect this build does not reproduce); avg = sum over active children / window volume;
                        max backward: first maximal child takes the gradient)
  refine             ops/Refine.cu:17-58                        fine takes floor(fine / factor)'s features if that voxel is active
Parity: these are plain definitions; the reference's tests for them (tests/unit/test_basic_ops.py max_pool / refine cases)
compare against dense torch pooling, which test_pool_refine_* in tests/test_gpu_parity.py restates for a dense block.
"""

from __future__ import annotations

import itertools

import numpy as np


def coarsened_ijk(ijk: np.ndarray, factor) -> np.ndarray:
    f = np.asarray(factor, dtype=np.int64)
    return np.unique(np.floor_divide(ijk.astype(np.int64), f), axis=0)


def refined_ijk(ijk: np.ndarray, factor) -> np.ndarray:
    f = np.asarray(factor, dtype=np.int64)
    cells = np.array(list(itertools.product(range(f[0]), range(f[1]), range(f[2]))), dtype=np.int64)
    return np.unique((ijk.astype(np.int64)[:, None, :] * f + cells[None]).reshape(-1, 3), axis=0)


def coarse_metadata(voxel_size, origin, factor):
    f, s, o = (np.asarray(v, dtype=np.float64) for v in (factor, voxel_size, origin))
    return f * s, (f - 1.0) * s * 0.5 + o


def fine_metadata(voxel_size, origin, factor):
    f, s, o = (np.asarray(v, dtype=np.float64) for v in (factor, voxel_size, origin))
    return s / f, o - (f - 1.0) * (s / f) * 0.5


def _rows(ijk, bidx):
    return {(int(b), int(i), int(j), int(k)): r for r, (b, (i, j, k)) in enumerate(zip(bidx, ijk))}


def pool(fine_ijk, fine_bidx, x, coarse_ijk, coarse_bidx, factor, stride, mode: str):
    """Returns (y [Nc, C], children [Nc, taps] of fine rows or -1)."""
    lut = _rows(fine_ijk, fine_bidx)
    f = [int(v) for v in factor]
    st = [int(s) if int(s) > 0 else fv for s, fv in zip(stride, f)]
    cells = list(itertools.product(range(f[0]), range(f[1]), range(f[2])))
    children = np.full((len(coarse_ijk), len(cells)), -1, dtype=np.int64)
    y = np.full((len(coarse_ijk), x.shape[1]), -np.inf if mode == "max" else 0.0, dtype=np.float64)
    for r, (b, c) in enumerate(zip(coarse_bidx, coarse_ijk)):
        for t, cell in enumerate(cells):
            key = (int(b), int(c[0]) * st[0] + cell[0], int(c[1]) * st[1] + cell[1], int(c[2]) * st[2] + cell[2])
            if key in lut:
                children[r, t] = lut[key]
                y[r] = np.maximum(y[r], x[lut[key]]) if mode == "max" else y[r] + x[lut[key]]
    if mode == "avg":
        y /= float(len(cells))
    y[(children < 0).all(axis=1)] = 0.0
    return y, children


def pool_backward(dy, x, children, n_fine, mode: str):
    dx = np.zeros((n_fine, x.shape[1]), dtype=np.float64)
    taps = children.shape[1]
    for r in range(children.shape[0]):
        live = [int(c) for c in children[r] if c >= 0]
        if not live:
            continue
        if mode == "avg":
            for c in live:
                dx[c] = dy[r] / taps
        else:
            vals = x[live]  # [children, C]
            arg = np.argmax(vals, axis=0)  # first maximum
            for ch in range(x.shape[1]):
                dx[live[arg[ch]], ch] = dy[r, ch]
    return dx


def refine(coarse_ijk, coarse_bidx, x, fine_ijk, fine_bidx, factor):
    lut = _rows(coarse_ijk, coarse_bidx)
    f = np.asarray(factor, dtype=np.int64)
    parent = np.full(len(fine_ijk), -1, dtype=np.int64)
    y = np.zeros((len(fine_ijk), x.shape[1]), dtype=np.float64)
    for r, (b, p) in enumerate(zip(fine_bidx, np.floor_divide(fine_ijk.astype(np.int64), f))):
        key = (int(b), int(p[0]), int(p[1]), int(p[2]))
        if key in lut:
            parent[r] = lut[key]
            y[r] = x[lut[key]]
    return y, parent


def refine_backward(dy, parent, n_coarse):
    dx = np.zeros((n_coarse, dy.shape[1]), dtype=np.float64)
    for r, p in enumerate(parent):
        if p >= 0:
            dx[p] += dy[r]
    return dx
