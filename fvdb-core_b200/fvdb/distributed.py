"""Multi-GPU plumbing for the convolution path: partition a batch BY GRID, all-reduce weight gradients.

The kernel map never crosses grids (reference GatherScatterDefault.cu:126,186-188), so each rank owns
whole grids and forward / dgrad need no collective; the training step's only exchange is a SUM
all-reduce of ``grad_weights`` (+ bias grads) over NCCL (gloo in the CPU tests).  One process per GPU.
"""

from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


def partition_grids_lpt(voxel_counts: Sequence[int], world_size: int) -> list[list[int]]:
    """Longest-processing-time bin packing of grid indices by voxel count; deterministic.

    Returns ``world_size`` lists of grid indices (each sorted ascending).  Every grid lands on exactly one rank.
    """
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    loads = [0] * world_size
    bins: list[list[int]] = [[] for _ in range(world_size)]
    order = sorted(range(len(voxel_counts)), key=lambda g: (-int(voxel_counts[g]), g))
    for g in order:
        r = min(range(world_size), key=lambda i: (loads[i], i))
        bins[r].append(g)
        loads[r] += int(voxel_counts[g])
    return [sorted(b) for b in bins]


def allreduce_gradients(parameters: Iterable[torch.nn.Parameter], group=None, bucket_bytes: int = 32 << 20) -> int:
    """SUM all-reduce of ``.grad`` over the process group, flattened into buckets (size chosen for launch
    latency, not link count: NVSwitch gives every peer full bandwidth).  Returns the number of collectives."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad for p in parameters if p.grad is not None]
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        offset = 0
        for g in bucket:
            g.copy_(flat[offset : offset + g.numel()].view_as(g))
            offset += g.numel()
        calls += 1
        bucket, size = [], 0

    by_dtype: dict[torch.dtype, list[torch.Tensor]] = {}
    for g in grads:
        by_dtype.setdefault(g.dtype, []).append(g)
    for group_grads in by_dtype.values():
        for g in group_grads:
            if size + g.numel() * g.element_size() > bucket_bytes:
                flush()
            bucket.append(g)
            size += g.numel() * g.element_size()
        flush()
    return calls
