"""Argument normalisation for the convolution path (subset of reference fvdb/types.py:1427 ``to_Vec3i``)."""

from __future__ import annotations

from enum import Enum
from typing import Sequence, Union

import numpy as np
import torch

NumericMaxRank1 = Union[int, float, Sequence[int], Sequence[float], np.ndarray, torch.Tensor]
NumericMaxRank2 = Union[int, float, Sequence, np.ndarray, torch.Tensor]
DeviceIdentifier = Union[str, torch.device]


class ValueConstraint(Enum):
    NONE = 0
    NON_NEGATIVE = 1
    POSITIVE = 2


def to_Vec3i(x: NumericMaxRank1, value_constraint: ValueConstraint = ValueConstraint.NONE) -> torch.Tensor:
    """Broadcast an int / 1- or 3-sequence / tensor to an int32 CPU tensor of shape ``(3,)``."""
    if isinstance(x, torch.Tensor):
        t = x.detach().cpu()
    else:
        t = torch.as_tensor(np.asarray(x))
    if t.dtype.is_floating_point or t.dtype == torch.bool:
        raise TypeError(f"expected integer values, got dtype {t.dtype}")
    if t.ndim > 1:
        raise ValueError(f"expected a scalar or rank-1 value, got shape {tuple(t.shape)}")
    t = t.reshape(-1).to(torch.int64)
    if t.numel() == 1:
        t = t.repeat(3)
    if t.numel() != 3:
        raise ValueError(f"expected 1 or 3 values, got {t.numel()}")
    if value_constraint is ValueConstraint.POSITIVE and bool((t <= 0).any()):
        raise ValueError(f"all values must be positive, got {t.tolist()}")
    if value_constraint is ValueConstraint.NON_NEGATIVE and bool((t < 0).any()):
        raise ValueError(f"all values must be non-negative, got {t.tolist()}")
    return t.to(torch.int32)


def to_Vec3fBatch(x: NumericMaxRank2, batch_size: int, name: str, positive: bool = False) -> torch.Tensor:
    """Broadcast to a float64 CPU tensor of shape ``(batch_size, 3)`` (voxel sizes / origins)."""
    t = x.detach().cpu().to(torch.float64) if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))
    if t.ndim == 0:
        t = t.reshape(1, 1).expand(batch_size, 3)
    elif t.ndim == 1:
        if t.numel() == 1:
            t = t.reshape(1, 1).expand(batch_size, 3)
        elif t.numel() == 3:
            t = t.reshape(1, 3).expand(batch_size, 3)
        else:
            raise ValueError(f"{name} must broadcast to ({batch_size}, 3), got shape {tuple(t.shape)}")
    elif t.ndim == 2:
        if t.shape[1] != 3 or t.shape[0] not in (1, batch_size):
            raise ValueError(f"{name} must broadcast to ({batch_size}, 3), got shape {tuple(t.shape)}")
        t = t.expand(batch_size, 3)
    else:
        raise ValueError(f"{name} must have rank <= 2")
    t = t.contiguous().clone()
    if positive and bool((t <= 0).any()):
        raise ValueError(f"{name} must be positive")
    return t
