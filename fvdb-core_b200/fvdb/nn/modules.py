"""``fvdb.nn.SparseConv3d`` / ``SparseConvTranspose3d`` (mirror of reference fvdb/nn/modules.py:264-435).

Same constructor arguments, parameter names, shapes and memory order, so ``state_dict``s interchange with
the reference: ``weight`` is ``[Cout, Cin, k0, k1, k2]`` viewed over ``(k2, k1, k0, Cin, Cout)`` memory
(modules.py:282-289), ``[Cout, Cin]`` when the kernel volume is 1; uniform init with
``1 / sqrt(Cin * K^3)`` (modules.py:304-309).
"""

from __future__ import annotations

import math

import torch
from torch import nn
from torch.profiler import record_function

from ..convolution_plan import ConvolutionPlan
from ..jagged_tensor import JaggedTensor
from ..types import NumericMaxRank1, ValueConstraint, to_Vec3i


class _SparseConv3dBase(nn.Module):
    _transposed = False

    def __init__(self, in_channels: int, out_channels: int, kernel_size: NumericMaxRank1 = 3, stride: NumericMaxRank1 = 1, bias: bool = True) -> None:
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = to_Vec3i(kernel_size, value_constraint=ValueConstraint.POSITIVE)
        self.stride = to_Vec3i(stride, value_constraint=ValueConstraint.POSITIVE)
        self.kernel_volume = int(torch.prod(self.kernel_size).item())
        if self.kernel_volume > 1:
            k0, k1, k2 = self.kernel_size.tolist()
            # tap-major, Cout-fastest storage: the engine's weight pre-pack reads it with unit stride
            self.weight = nn.Parameter(torch.zeros(k2, k1, k0, in_channels, out_channels).permute(4, 3, 2, 1, 0))
        else:
            self.weight = nn.Parameter(torch.zeros(out_channels, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def extra_repr(self) -> str:
        text = f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}"
        return text + (", bias=False" if self.bias is None else "")

    def reset_parameters(self) -> None:
        bound = 1 / math.sqrt(self.in_channels * self.kernel_volume)
        self.weight.data.uniform_(-bound, bound)
        if self.bias is not None:
            self.bias.data.uniform_(-bound, bound)

    def forward(self, data: JaggedTensor, plan: ConvolutionPlan) -> JaggedTensor:
        name = "SparseConvTranspose3d" if self._transposed else "SparseConv3d"
        with record_function(repr(self)):  # reference: @_trace_fvdb_nn_forward, modules.py:15-31
            if not plan.valid_usage(self.in_channels, self.out_channels, self.kernel_size, self.stride, transposed=self._transposed):
                raise ValueError(
                    f"Convolution plan used with a {name} module that had "
                    "mismatched input/output channels, kernel size, or stride, or transposition"
                )
            assert isinstance(data, JaggedTensor), "Input data must be a JaggedTensor"
            # the bias add of the reference (modules.py:370-371) is fused into the convolution kernel's epilogue
            return plan.execute(data, self.weight, self.bias)


class SparseConv3d(_SparseConv3dBase):
    """Sparse 3-D convolution over a JaggedTensor according to a (non-transposed) ConvolutionPlan."""

    _transposed = False


class SparseConvTranspose3d(_SparseConv3dBase):
    """Sparse 3-D transposed convolution according to a transposed ConvolutionPlan."""

    _transposed = True
