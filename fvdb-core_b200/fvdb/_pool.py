"""Pooling and nearest-neighbour refinement between a fine and a coarse GridBatch (reference: GridBatch.max_pool /
avg_pool / refine, fvdb/grid_batch.py:463-490,1139-1164,1474-1499 over ops/MaxPool.cu, ops/AvgPool.cu, ops/Refine.cu).

The window children of every coarse voxel are looked up once, with the index-grid lookup kernel, into a child table
``idx[N_coarse, taps]``; the data movement is done by the streaming kernels of csrc/pool.cu.  Window of coarse voxel
``c``: fine voxels ``stride * c + [0, pool_factor)^3`` (MaxPool.cu:38-50 -- no centring pad, unlike a convolution).
"""

from __future__ import annotations

import torch

from . import _fvdb_cpp, _lib
from ._lib import check, lib

_CODES = {torch.float16: _lib.FVC_F16, torch.bfloat16: _lib.FVC_BF16, torch.float32: _lib.FVC_F32}
POOL_MAX, POOL_SUM = 0, 1


def _check_rows(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{what}: this build runs on CUDA (sm_100a) only; got device {x.device}")
    if x.dim() != 2:
        raise RuntimeError(f"{what}: data must be [N, C], got {tuple(x.shape)}")


def _native(x: torch.Tensor) -> bool:
    """Row kernels serve f16 / bf16 / f32 with whole 16-byte channel vectors; other shapes (e.g. the 3- or 5-channel ends
    of a network, fp64) go through torch index ops on the same child tables."""
    return x.dtype in _CODES and x.shape[1] % (4 if x.dtype == torch.float32 else 8) == 0 and x.shape[1] > 0


def _pool_torch(x, idx, mode, scale):
    live = idx >= 0
    rows = x[idx.clamp_min(0).long()]  # [n_out, taps, C]
    if mode == POOL_MAX:
        y = rows.masked_fill(~live[:, :, None], float("-inf")).amax(dim=1)
        return torch.where(live.any(dim=1, keepdim=True), y, torch.zeros_like(y))
    return (rows * live[:, :, None]).sum(dim=1) * scale


def window_children(fine, coarse, factor: list[int], stride: list[int]) -> torch.Tensor:
    """``idx[N_coarse, f0*f1*f2]`` int32: batch-cumulative fine row of every window cell of every coarse voxel, -1 if inactive."""
    dev = coarse.device
    f0, f1, f2 = factor
    taps = torch.stack(torch.meshgrid(torch.arange(f0), torch.arange(f1), torch.arange(f2), indexing="ij"), dim=-1).reshape(-1, 3).to(dev, torch.int32)
    base = coarse.data.ijk * torch.tensor(stride, dtype=torch.int32, device=dev)
    queries = (base[:, None, :] + taps[None, :, :]).reshape(-1, 3).contiguous()
    jidx = coarse.data.jidx.repeat_interleave(taps.shape[0]) if coarse.grid_count > 1 else None
    idx = _fvdb_cpp.ijk_to_index(fine.data, queries, jidx, cumulative=True)
    return idx.to(torch.int32).reshape(coarse.total_voxels, taps.shape[0]).contiguous()


def parent_rows(coarse, fine, factor: list[int]) -> torch.Tensor:
    """``idx[N_fine]`` int32: batch-cumulative coarse row of ``floor(fine_ijk / factor)``, -1 if inactive (Refine.cu:43-52)."""
    f = torch.tensor(factor, dtype=torch.int32, device=fine.device)
    parents = torch.div(fine.data.ijk, f, rounding_mode="floor").to(torch.int32).contiguous()
    idx = _fvdb_cpp.ijk_to_index(coarse.data, parents, fine.data.jidx if fine.grid_count > 1 else None, cumulative=True)
    return idx.to(torch.int32).contiguous()


def _stream(x):
    return torch.cuda.current_stream(x.device).cuda_stream


class PoolRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, mode, scale):  # type: ignore[override]
        x = x.contiguous()
        n_out, taps = idx.shape
        y = torch.empty((n_out, x.shape[1]), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.fvc_pool_rows(x.data_ptr(), idx.data_ptr(), n_out, taps, x.shape[1], _CODES[x.dtype], mode, float(scale), y.data_ptr(), _stream(x)))
        ctx.save_for_backward(x, idx)
        ctx.cfg = (mode, float(scale))
        return y

    @staticmethod
    def backward(ctx, grad_output):  # type: ignore[override]
        x, idx = ctx.saved_tensors
        mode, scale = ctx.cfg
        dy = grad_output.contiguous()
        dx = torch.zeros_like(x)
        with torch.cuda.device(x.device):
            check(lib.fvc_pool_rows_backward(dy.data_ptr(), x.data_ptr(), idx.data_ptr(), idx.shape[0], idx.shape[1], x.shape[1], _CODES[x.dtype], mode, scale,
                                             dx.data_ptr(), _stream(x)))
        return dx, None, None, None


class RefineRowsFn(torch.autograd.Function):
    """fine[r] = coarse[parent[r]]; the backward is the sum over each coarse voxel's children (same pooling kernel)."""

    @staticmethod
    def forward(ctx, x, parent, children):  # type: ignore[override]
        x = x.contiguous()
        y = torch.empty((parent.shape[0], x.shape[1]), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.fvc_gather_rows(x.data_ptr(), parent.data_ptr(), parent.shape[0], x.shape[1], _CODES[x.dtype], y.data_ptr(), _stream(x)))
        ctx.save_for_backward(children)
        return y

    @staticmethod
    def backward(ctx, grad_output):  # type: ignore[override]
        (children,) = ctx.saved_tensors
        dy = grad_output.contiguous()
        dx = torch.empty((children.shape[0], dy.shape[1]), dtype=dy.dtype, device=dy.device)
        with torch.cuda.device(dy.device):
            check(lib.fvc_pool_rows(dy.data_ptr(), children.data_ptr(), children.shape[0], children.shape[1], dy.shape[1], _CODES[dy.dtype], POOL_SUM, 1.0,
                                    dx.data_ptr(), _stream(dy)))
        return dx, None, None


def pool_rows(x: torch.Tensor, idx: torch.Tensor, mode: int, scale: float) -> torch.Tensor:
    _check_rows(x, "pool")
    return PoolRowsFn.apply(x, idx, mode, scale) if _native(x) else _pool_torch(x, idx, mode, scale)


def refine_rows(x: torch.Tensor, parent: torch.Tensor, children: torch.Tensor) -> torch.Tensor:
    _check_rows(x, "refine")
    if _native(x):
        return RefineRowsFn.apply(x, parent, children)
    return x[parent.clamp_min(0).long()] * (parent >= 0)[:, None]
