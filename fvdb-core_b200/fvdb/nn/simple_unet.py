"""Residual sparse U-Net over the ConvolutionPlan path (mirror of reference fvdb/nn/simple_unet.py).

Same class names, constructor arguments, sub-module attribute names (``state_dict`` keys interchange with the reference)
and data flow: pad (conv onto the dilated grid) -> recursive down / up levels (max-pool, 1x1x1 fan-out, residual blocks,
1x1x1 fan-in, nearest refinement, additive skip) -> unpad (transposed conv back onto the input grid).  Every convolution
goes through ``ConvolutionPlan``; BatchNorm + ReLU pairs run as one fused pass (csrc/norm.cu) -- the same function as the
reference's separate ``BatchNorm`` then ``fvdb.torch_jagged.relu``.
"""

from __future__ import annotations

import torch
from torch import nn

from ..convolution_plan import ConvolutionPlan
from ..grid_batch import GridBatch
from ..jagged_tensor import JaggedTensor
from ..types import NumericMaxRank1
from .modules import BatchNorm, MaxPool, SparseConv3d, SparseConvTranspose3d


def _same_grid_plan(kernel_size, grid: GridBatch) -> ConvolutionPlan:
    return ConvolutionPlan.from_grid_batch(kernel_size=kernel_size, stride=1, source_grid=grid, target_grid=grid)


class SimpleUNetBasicBlock(nn.Module):
    """conv (no bias) -> batch norm -> ReLU  (simple_unet.py:44-99)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: NumericMaxRank1 = 3, momentum: float = 0.1) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size, self.momentum = in_channels, out_channels, kernel_size, momentum
        self.conv = SparseConv3d(in_channels, out_channels, kernel_size=kernel_size, stride=1, bias=False)
        self.batch_norm = BatchNorm(out_channels, momentum=momentum, activation="relu")

    def extra_repr(self) -> str:
        return f"in_channels={self.in_channels}, out_channels={self.out_channels}, kernel_size={self.kernel_size}, momentum={self.momentum}"

    def reset_parameters(self) -> None:
        self.conv.reset_parameters()
        self.batch_norm.reset_parameters()

    def forward(self, data: JaggedTensor, plan: ConvolutionPlan) -> JaggedTensor:
        return self.batch_norm(self.conv(data, plan), plan.target_grid_batch)


class SimpleUNetConvBlock(nn.Module):
    """``layer_count`` basic blocks with an additive residual and a final ReLU (simple_unet.py:103-190)."""

    def __init__(self, in_channels: int, mid_channels: int, out_channels: int, kernel_size: NumericMaxRank1 = 3, layer_count: int = 2,
                 momentum: float = 0.1) -> None:
        super().__init__()
        self.in_channels, self.mid_channels, self.out_channels = in_channels, mid_channels, out_channels
        self.kernel_size, self.layer_count, self.momentum = kernel_size, layer_count, momentum
        widths = [in_channels] + [mid_channels] * (layer_count - 1) + [out_channels]
        self.blocks = nn.ModuleList(SimpleUNetBasicBlock(widths[i], widths[i + 1], kernel_size, momentum) for i in range(layer_count))

    def extra_repr(self) -> str:
        return (f"in_channels={self.in_channels}, mid_channels={self.mid_channels}, out_channels={self.out_channels}, "
                f"kernel_size={self.kernel_size}, layer_count={self.layer_count}, momentum={self.momentum}")

    def reset_parameters(self) -> None:
        for block in self.blocks:
            block.reset_parameters()

    def forward(self, data: JaggedTensor, plan: ConvolutionPlan) -> JaggedTensor:
        if not plan.has_fixed_topology:
            raise ValueError("Convolution plan must have fixed topology for repeated conv blocks.")
        residual = data
        for block in self.blocks:
            data = block(data, plan)
        out = data + residual
        return out.jagged_like(torch.relu(out.jdata))


class SimpleUNetDown(nn.Module):
    """max-pool by 2, then a 1x1x1 channel fan-out and batch norm on the coarse grid (simple_unet.py:194-243)."""

    def __init__(self, in_channels: int, out_channels: int, momentum: float = 0.1) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.momentum = in_channels, out_channels, momentum
        self.max_pool = MaxPool(kernel_size=2)
        self.channel_fan_out = SparseConv3d(in_channels, out_channels, kernel_size=1, stride=1, bias=False)
        self.batch_norm = BatchNorm(out_channels, momentum=momentum)

    def extra_repr(self) -> str:
        return f"in_channels={self.in_channels}, out_channels={self.out_channels}, momentum={self.momentum}"

    def reset_parameters(self) -> None:
        self.channel_fan_out.reset_parameters()
        self.batch_norm.reset_parameters()

    def forward(self, data: JaggedTensor, fine_grid: GridBatch, coarse_grid: GridBatch) -> JaggedTensor:
        data, _ = self.max_pool(data, fine_grid, coarse_grid)
        return self.batch_norm(self.channel_fan_out(data, _same_grid_plan(1, coarse_grid)), coarse_grid)


class SimpleUNetUp(nn.Module):
    """1x1x1 channel fan-in and batch norm on the coarse grid, then nearest refinement by 2 (simple_unet.py:246-293)."""

    def __init__(self, in_channels: int, out_channels: int, momentum: float = 0.1) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.momentum = in_channels, out_channels, momentum
        self.channel_fan_in = SparseConv3d(in_channels, out_channels, kernel_size=1, stride=1, bias=False)
        self.batch_norm = BatchNorm(out_channels, momentum=momentum)

    def extra_repr(self) -> str:
        return f"in_channels={self.in_channels}, out_channels={self.out_channels}, momentum={self.momentum}"

    def reset_parameters(self) -> None:
        self.channel_fan_in.reset_parameters()
        self.batch_norm.reset_parameters()

    def forward(self, data: JaggedTensor, coarse_grid: GridBatch, fine_grid: GridBatch) -> JaggedTensor:
        data = self.batch_norm(self.channel_fan_in(data, _same_grid_plan(1, coarse_grid)), coarse_grid)
        return coarse_grid.refine(subdiv_factor=2, data=data, fine_grid=fine_grid)[0]


class SimpleUNetBottleneck(nn.Module):
    """One residual block at the coarsest resolution (simple_unet.py:297-341)."""

    def __init__(self, channels: int, kernel_size: NumericMaxRank1 = 3, layer_count: int = 2, momentum: float = 0.1) -> None:
        super().__init__()
        self.channels, self.kernel_size, self.layer_count, self.momentum = channels, kernel_size, layer_count, momentum
        self.block = SimpleUNetConvBlock(channels, channels, channels, kernel_size, layer_count, momentum)

    def extra_repr(self) -> str:
        return f"channels={self.channels}, kernel_size={self.kernel_size}, layer_count={self.layer_count}, momentum={self.momentum}"

    def reset_parameters(self) -> None:
        self.block.reset_parameters()

    def forward(self, data: JaggedTensor, grid: GridBatch) -> JaggedTensor:
        return self.block(data, _same_grid_plan(self.block.kernel_size, grid))


class SimpleUNetDownUp(nn.Module):
    """One resolution level: block, down, (inner level or bottleneck), up, additive skip, block (simple_unet.py:344-450)."""

    def __init__(self, in_channels: int, channel_growth_rate: int, kernel_size: NumericMaxRank1 = 3, downup_layer_count: int = 4,
                 block_layer_count: int = 2, momentum: float = 0.1):
        super().__init__()
        self.in_channels, self.channel_growth_rate, self.kernel_size = in_channels, channel_growth_rate, kernel_size
        self.downup_layer_count, self.block_layer_count, self.momentum = downup_layer_count, block_layer_count, momentum
        coarse_channels = in_channels * channel_growth_rate
        self.conv_in = SimpleUNetConvBlock(in_channels, in_channels, in_channels, kernel_size, block_layer_count, momentum)
        self.down = SimpleUNetDown(in_channels, coarse_channels, momentum)
        if downup_layer_count <= 1:
            self.inner = SimpleUNetBottleneck(coarse_channels, kernel_size, block_layer_count, momentum)
        else:
            self.inner = SimpleUNetDownUp(coarse_channels, channel_growth_rate, kernel_size, downup_layer_count - 1, block_layer_count, momentum)
        self.up = SimpleUNetUp(coarse_channels, in_channels, momentum)
        self.conv_out = SimpleUNetConvBlock(in_channels, in_channels, in_channels, kernel_size, block_layer_count, momentum)

    def extra_repr(self) -> str:
        return (f"in_channels={self.in_channels}, channel_growth_rate={self.channel_growth_rate}, kernel_size={self.kernel_size}, "
                f"downup_layer_count={self.downup_layer_count}, block_layer_count={self.block_layer_count}, momentum={self.momentum}")

    def reset_parameters(self) -> None:
        for part in (self.conv_in, self.down, self.inner, self.up, self.conv_out):
            part.reset_parameters()

    def forward(self, data: JaggedTensor, fine_grid: GridBatch) -> JaggedTensor:
        # the coarse level lives on the block-centroid grid, dilated so that its stride-1 convolutions see their whole support
        coarse_grid = fine_grid.coarsened_grid(coarsening_factor=2).conv_grid(kernel_size=self.kernel_size, stride=1)
        plan = _same_grid_plan(self.kernel_size, fine_grid)
        skip = data
        data = self.conv_in(data, plan)
        data = self.down(data, fine_grid, coarse_grid)
        data = self.inner(data, coarse_grid)
        data = self.up(data, coarse_grid, fine_grid)
        return self.conv_out(data + skip, plan)


class SimpleUNetPad(nn.Module):
    """Convolution from the input grid onto its dilation, input channels -> base channels (simple_unet.py:453-504)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: NumericMaxRank1 = 3, momentum: float = 0.1) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size, self.momentum = in_channels, out_channels, kernel_size, momentum
        self.conv = SparseConv3d(in_channels, out_channels, kernel_size=kernel_size, stride=1, bias=False)
        self.batch_norm = BatchNorm(out_channels, momentum=momentum)

    def extra_repr(self) -> str:
        return f"in_channels={self.in_channels}, out_channels={self.out_channels}, kernel_size={self.kernel_size}, momentum={self.momentum}"

    def reset_parameters(self) -> None:
        self.conv.reset_parameters()
        self.batch_norm.reset_parameters()

    def create_padded_grid(self, grid: GridBatch) -> GridBatch:
        return grid.conv_grid(kernel_size=self.kernel_size, stride=1)

    def forward(self, data: JaggedTensor, grid: GridBatch, padded_grid: GridBatch) -> JaggedTensor:
        plan = ConvolutionPlan.from_grid_batch(kernel_size=self.kernel_size, stride=1, source_grid=grid, target_grid=padded_grid)
        return self.batch_norm(self.conv(data, plan), padded_grid)


class SimpleUNetUnpad(nn.Module):
    """Transposed convolution from the padded grid back onto the input grid, base channels -> output channels (:507-553)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: NumericMaxRank1 = 3) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.deconv = SparseConvTranspose3d(in_channels, out_channels, kernel_size=kernel_size, stride=1, bias=False)

    def extra_repr(self) -> str:
        return f"in_channels={self.in_channels}, out_channels={self.out_channels}, kernel_size={self.kernel_size}"

    def reset_parameters(self) -> None:
        self.deconv.reset_parameters()

    def forward(self, data: JaggedTensor, padded_grid: GridBatch, grid: GridBatch) -> JaggedTensor:
        plan = ConvolutionPlan.from_grid_batch_transposed(kernel_size=self.kernel_size, stride=1, source_grid=padded_grid, target_grid=grid)
        return self.deconv(data, plan)


class SimpleUNet(nn.Module):
    """pad -> down/up levels -> unpad (simple_unet.py:555-640)."""

    def __init__(self, in_channels: int, base_channels: int, out_channels: int, channel_growth_rate: int, kernel_size: NumericMaxRank1 = 3,
                 downup_layer_count: int = 4, block_layer_count: int = 2, momentum: float = 0.1):
        super().__init__()
        self.in_channels, self.base_channels, self.out_channels = in_channels, base_channels, out_channels
        self.channel_growth_rate, self.kernel_size = channel_growth_rate, kernel_size
        self.downup_layer_count, self.block_layer_count, self.momentum = downup_layer_count, block_layer_count, momentum
        self.pad = SimpleUNetPad(in_channels, base_channels, kernel_size, momentum)
        self.downup = SimpleUNetDownUp(base_channels, channel_growth_rate, kernel_size, downup_layer_count, block_layer_count, momentum)
        self.unpad = SimpleUNetUnpad(base_channels, out_channels, kernel_size)

    def extra_repr(self) -> str:
        return (f"in_channels={self.in_channels}, base_channels={self.base_channels}, out_channels={self.out_channels}, "
                f"channel_growth_rate={self.channel_growth_rate}, kernel_size={self.kernel_size}, downup_layer_count={self.downup_layer_count}, "
                f"block_layer_count={self.block_layer_count}, momentum={self.momentum}")

    def reset_parameters(self) -> None:
        for part in (self.pad, self.downup, self.unpad):
            part.reset_parameters()

    def forward(self, data: JaggedTensor, grid: GridBatch) -> JaggedTensor:
        padded_grid = self.pad.create_padded_grid(grid)
        data = self.pad(data, grid, padded_grid)
        data = self.downup(data, padded_grid)
        return self.unpad(data, padded_grid, grid)
