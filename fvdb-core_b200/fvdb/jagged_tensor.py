"""Minimal ``JaggedTensor``: the ragged batch container the convolution path exchanges.

Mirror of the subset of reference ``fvdb.JaggedTensor`` (fvdb/jagged_tensor.py, src/fvdb/JaggedTensor.h:16-22)
that ``ConvolutionPlan`` / ``fvdb.nn`` touch: ``jdata [N, ...]``, ``joffsets int64 [B+1]``, ``jidx int32 [N]``,
``jagged_like``, ``unbind``, indexing by batch item and elementwise arithmetic on ``jdata``.  It is plain
torch on whatever device ``jdata`` lives on; the rest of the reference's jagged op surface is out of scope
(SURVEY.md section 2, row 13).
"""

from __future__ import annotations

from typing import Any, Iterator, Sequence

import torch

JIdxType = torch.int32
JOffsetsType = torch.int64


class JaggedTensor:
    def __init__(self, tensors: "torch.Tensor | Sequence[torch.Tensor] | JaggedTensor | None" = None, *, _data=None, _offsets=None, _jidx=None):
        if _data is not None:
            self._jdata, self._joffsets, self._jidx = _data, _offsets, _jidx
            return
        if isinstance(tensors, JaggedTensor):
            self._jdata, self._joffsets, self._jidx = tensors._jdata, tensors._joffsets, tensors._jidx
            return
        if isinstance(tensors, torch.Tensor):
            tensors = [tensors]
        if tensors is None or len(tensors) == 0:
            raise ValueError("JaggedTensor needs a tensor or a non-empty list of tensors")
        if any(isinstance(t, (list, tuple)) for t in tensors):
            raise ValueError("lists of lists of tensors are not supported by this build")
        first = tensors[0]
        for t in tensors:
            if not isinstance(t, torch.Tensor):
                raise TypeError("JaggedTensor elements must be torch.Tensors")
            if t.device != first.device or t.dtype != first.dtype or t.shape[1:] != first.shape[1:]:
                raise ValueError("all tensors in a JaggedTensor must share device, dtype and trailing shape")
        self._jdata = torch.cat(list(tensors), dim=0) if len(tensors) > 1 else first
        counts = torch.tensor([0] + [int(t.shape[0]) for t in tensors], dtype=JOffsetsType)
        self._joffsets = torch.cumsum(counts, 0).to(first.device)
        self._jidx = None

    # ---- constructors -----------------------------------------------------------------------
    @classmethod
    def from_tensor(cls, data: torch.Tensor) -> "JaggedTensor":
        return cls(data)

    @classmethod
    def from_list_of_tensors(cls, tensors: Sequence[torch.Tensor]) -> "JaggedTensor":
        return cls(list(tensors))

    @classmethod
    def from_data_and_offsets(cls, data: torch.Tensor, offsets: torch.Tensor) -> "JaggedTensor":
        offsets = offsets.to(device=data.device, dtype=JOffsetsType)
        if offsets.ndim != 1 or offsets.numel() < 1:
            raise ValueError("offsets must be a 1-D tensor with at least one entry")
        return cls(_data=data, _offsets=offsets, _jidx=None)

    @classmethod
    def from_data_and_indices(cls, data: torch.Tensor, indices: torch.Tensor, num_tensors: int) -> "JaggedTensor":
        counts = torch.bincount(indices.to(torch.int64), minlength=num_tensors)
        offsets = torch.zeros(num_tensors + 1, dtype=JOffsetsType, device=data.device)
        offsets[1:] = torch.cumsum(counts, 0)
        return cls(_data=data, _offsets=offsets, _jidx=indices.to(JIdxType))

    # ---- structure --------------------------------------------------------------------------
    @property
    def jdata(self) -> torch.Tensor:
        return self._jdata

    @jdata.setter
    def jdata(self, value: torch.Tensor) -> None:
        if value.shape[0] != self._jdata.shape[0]:
            raise ValueError("new jdata must keep the leading (voxel) dimension")
        self._jdata = value

    @property
    def joffsets(self) -> torch.Tensor:
        return self._joffsets

    @property
    def jidx(self) -> torch.Tensor:
        if self._jidx is None:
            counts = (self._joffsets[1:] - self._joffsets[:-1]).to(torch.int64)
            self._jidx = torch.repeat_interleave(
                torch.arange(counts.numel(), device=self._jdata.device, dtype=JIdxType), counts, output_size=int(self._jdata.shape[0])
            )
        return self._jidx

    @property
    def num_tensors(self) -> int:
        return int(self._joffsets.numel() - 1)

    def __len__(self) -> int:
        return self.num_tensors

    @property
    def device(self) -> torch.device:
        return self._jdata.device

    @property
    def dtype(self) -> torch.dtype:
        return self._jdata.dtype

    @property
    def is_cuda(self) -> bool:
        return self._jdata.is_cuda

    @property
    def is_cpu(self) -> bool:
        return self._jdata.is_cpu

    @property
    def rshape(self) -> tuple[int, ...]:
        return tuple(self._jdata.shape)

    @property
    def eshape(self) -> list[int]:
        return list(self._jdata.shape[1:])

    @property
    def lshape(self) -> list[int]:
        return (self._joffsets[1:] - self._joffsets[:-1]).tolist()

    @property
    def requires_grad(self) -> bool:
        return self._jdata.requires_grad

    def requires_grad_(self, requires_grad: bool = True) -> "JaggedTensor":
        self._jdata.requires_grad_(requires_grad)
        return self

    def jagged_like(self, data: torch.Tensor) -> "JaggedTensor":
        """A JaggedTensor with this one's structure and new data (reference jagged_tensor.py:860)."""
        if data.shape[0] != self._jdata.shape[0]:
            raise ValueError(f"data has {data.shape[0]} rows, expected {self._jdata.shape[0]}")
        return JaggedTensor(_data=data, _offsets=self._joffsets.to(data.device), _jidx=None if self._jidx is None else self._jidx.to(data.device))

    def unbind(self) -> list[torch.Tensor]:
        offsets = self._joffsets.tolist()
        return [self._jdata[offsets[i] : offsets[i + 1]] for i in range(len(offsets) - 1)]

    def __getitem__(self, index: Any) -> "JaggedTensor":
        if isinstance(index, int):
            n = self.num_tensors
            if index < -n or index >= n:
                raise IndexError(f"batch index {index} out of range for {n} tensors")
            index %= n
            lo, hi = int(self._joffsets[index]), int(self._joffsets[index + 1])
            return JaggedTensor(self._jdata[lo:hi])
        if isinstance(index, slice):
            parts = self.unbind()[index]
            if not parts:
                raise IndexError("empty JaggedTensor slice")
            return JaggedTensor(parts)
        raise TypeError("JaggedTensor supports int and slice indexing in this build")

    def __iter__(self) -> Iterator["JaggedTensor"]:
        for i in range(self.num_tensors):
            yield self[i]

    # ---- movement / dtype -------------------------------------------------------------------
    def to(self, device_or_dtype) -> "JaggedTensor":
        data = self._jdata.to(device_or_dtype)
        return JaggedTensor(_data=data, _offsets=self._joffsets.to(data.device), _jidx=None if self._jidx is None else self._jidx.to(data.device))

    def cpu(self) -> "JaggedTensor":
        return self.to("cpu")

    def cuda(self) -> "JaggedTensor":
        return self.to("cuda")

    def type(self, dtype: torch.dtype) -> "JaggedTensor":
        return self.jagged_like(self._jdata.to(dtype))

    def float(self) -> "JaggedTensor":
        return self.type(torch.float32)

    def double(self) -> "JaggedTensor":
        return self.type(torch.float64)

    def detach(self) -> "JaggedTensor":
        return self.jagged_like(self._jdata.detach())

    def clone(self) -> "JaggedTensor":
        return self.jagged_like(self._jdata.clone())

    # ---- elementwise arithmetic on jdata ----------------------------------------------------
    def _binary(self, other, op) -> "JaggedTensor":
        rhs = other._jdata if isinstance(other, JaggedTensor) else other
        return self.jagged_like(op(self._jdata, rhs))

    def __add__(self, other):
        return self._binary(other, torch.add)

    def __sub__(self, other):
        return self._binary(other, torch.sub)

    def __mul__(self, other):
        return self._binary(other, torch.mul)

    def __truediv__(self, other):
        return self._binary(other, torch.div)

    def __neg__(self):
        return self.jagged_like(-self._jdata)

    def __repr__(self) -> str:
        return f"JaggedTensor(num_tensors={self.num_tensors}, rshape={self.rshape}, dtype={self.dtype}, device={self.device})"
