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

from .. import _norm
from ..convolution_plan import ConvolutionPlan
from ..grid_batch import GridBatch
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


class BatchNorm(nn.BatchNorm1d):
    """Batch normalisation over the voxels of a JaggedTensor (mirror of reference fvdb/nn/modules.py:484-521: same
    constructor, parameters / buffers and ``forward(data, grid)``), executed by the streaming kernels of csrc/norm.cu.

    ``activation="relu"`` (extension, keyword only) fuses the ReLU that follows the norm in every block of the
    reference's networks (fvdb/nn/simple_unet.py) into the same pass, forward and backward."""

    _sync = False

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: "float | None" = 0.1, affine: bool = True,
                 track_running_stats: bool = True, device=None, dtype=None, *, activation: "str | None" = None, process_group=None) -> None:
        super().__init__(num_features, eps, momentum, affine, track_running_stats, device=device, dtype=dtype)
        if activation not in (None, "relu"):
            raise ValueError("activation must be None or 'relu'")
        self.activation, self.process_group = activation, process_group

    def _rows(self, x: torch.Tensor, conv_stats=None) -> torch.Tensor:
        training = self.training or self.running_mean is None
        momentum = 0.0 if self.momentum is None else self.momentum
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if self.momentum is None:  # cumulative moving average
                momentum = 1.0 / float(self.num_batches_tracked)
        self._effective_momentum = momentum if (training and self.track_running_stats) else 0.0
        group = None
        if self._sync and training:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
                group = self.process_group if self.process_group is not None else dist.group.WORLD
        return _norm.batch_norm_rows(x, self.weight, self.bias, self.running_mean if self.track_running_stats else None,
                                     self.running_var if self.track_running_stats else None, training, momentum, self.eps,
                                     relu=self.activation == "relu", group=group, conv_stats=conv_stats)

    def forward(self, data: JaggedTensor, grid=None) -> JaggedTensor:  # type: ignore[override]
        with record_function(repr(self)):
            num_channels = data.jdata.size(1)
            assert num_channels == self.num_features, "Input feature should have the same number of channels as BatchNorm"
            out = self._rows(data.jdata)
            return grid.jagged_like(out) if grid is not None else data.jagged_like(out)


class GroupNorm(nn.GroupNorm):
    """Group normalisation per grid of the batch (mirror of reference fvdb/nn/modules.py:438-480, which loops over the grids
    and calls torch's GroupNorm on each ``[1, C, N_b]`` slab).  Same statistics -- per (grid, group) over that grid's voxels
    and the group's channels -- computed for the whole batch at once with segment sums over ``jidx`` (torch ops, autograd)."""

    def forward(self, data: JaggedTensor, grid: GridBatch) -> JaggedTensor:  # type: ignore[override]
        with record_function(repr(self)):
            x = data.jdata
            n, c = x.shape
            assert c == self.num_channels, "Input feature should have the same number of channels as GroupNorm"
            groups, per = self.num_groups, c // self.num_groups
            batches = grid.grid_count
            owner = grid.jidx.long() if batches > 1 else torch.zeros(n, dtype=torch.long, device=x.device)
            work = torch.float64 if x.dtype == torch.float64 else torch.float32  # fp64 inputs keep the reference's fp64 accuracy
            xg = x.to(work).reshape(n, groups, per)
            count = torch.zeros(batches, device=x.device, dtype=work).index_add_(0, owner, torch.ones(n, device=x.device, dtype=work)).clamp_min(1.0) * per
            mean = torch.zeros((batches, groups), device=x.device, dtype=work).index_add_(0, owner, xg.sum(-1)) / count[:, None]
            centred = xg - mean[owner][:, :, None]
            var = torch.zeros((batches, groups), device=x.device, dtype=work).index_add_(0, owner, centred.square().sum(-1)) / count[:, None]
            out = (centred * torch.rsqrt(var + self.eps)[owner][:, :, None]).reshape(n, c)
            if self.affine:
                out = out * self.weight.to(work) + self.bias.to(work)
            return grid.jagged_like(out.to(x.dtype))


class SyncBatchNorm(BatchNorm):
    """BatchNorm with statistics over every process of ``process_group`` (mirror of reference modules.py:524-580): the
    per-rank (count, mean, M2) are all-gathered and merged, the two backward sums all-reduced (2*C floats each)."""

    _sync = True

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: "float | None" = 0.1, affine: bool = True,
                 track_running_stats: bool = True, process_group=None, device=None, dtype=None, *, activation: "str | None" = None) -> None:
        super().__init__(num_features, eps, momentum, affine, track_running_stats, device=device, dtype=dtype, activation=activation,
                         process_group=process_group)

    @classmethod
    def convert_sync_batchnorm(cls, module: nn.Module, process_group=None) -> nn.Module:
        """Replace every fvdb.nn.BatchNorm in ``module`` by a SyncBatchNorm sharing its parameters and buffers."""
        out = module
        if isinstance(module, BatchNorm) and not isinstance(module, SyncBatchNorm):
            out = cls(module.num_features, module.eps, module.momentum, module.affine, module.track_running_stats, process_group,
                      activation=module.activation)
            if module.affine:
                out.weight, out.bias = module.weight, module.bias
            out.running_mean, out.running_var, out.num_batches_tracked = module.running_mean, module.running_var, module.num_batches_tracked
            out.training = module.training
        for name, child in module.named_children():
            out.add_module(name, cls.convert_sync_batchnorm(child, process_group))
        return out


class _ZeroGradFor(torch.autograd.Function):
    """Identity on ``rows`` that reports an exactly-zero gradient for ``param`` (a parameter whose effect cancels downstream)."""

    @staticmethod
    def forward(ctx, rows, param):  # type: ignore[override]
        ctx.shape, ctx.dtype, ctx.device = param.shape, param.dtype, param.device
        return rows.view_as(rows)

    @staticmethod
    def backward(ctx, grad):  # type: ignore[override]
        return grad, torch.zeros(ctx.shape, dtype=ctx.dtype, device=ctx.device)


def conv_bn_act(conv: _SparseConv3dBase, norm: BatchNorm, data: JaggedTensor, plan: ConvolutionPlan, residual: "JaggedTensor | None" = None,
                final_relu: bool = False) -> JaggedTensor:
    """``norm(conv(data, plan)) + residual`` (then a ReLU if ``final_relu``: the tail of the reference's residual block,
    fvdb/nn/simple_unet.py:182-188) -- the conv -> BatchNorm -> ReLU block of the reference's networks
    (fvdb/nn/modules.py:484-521, fvdb/nn/simple_unet.py:233-243) with the passes over ``[N, C]`` folded into the GEMM epilogue:

    * training: the convolution epilogue also writes per-block column sums of its output, BatchNorm takes its batch
      statistics from them (no statistics pass) and applies normalisation + activation in one streaming pass; gradients
      flow through the ordinary backward of both modules;
    * inference (no grad, running statistics): BatchNorm folds into a per-channel scale / shift, and scale, shift, the
      residual add and the ReLU all run in the convolution kernel's epilogue -- ONE kernel for the whole block.

    Falls back to the separate modules wherever the fused epilogue does not exist (CUDA-core path, K = S = 1 matmul plans)."""
    name = "SparseConvTranspose3d" if conv._transposed else "SparseConv3d"
    if not plan.valid_usage(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, transposed=conv._transposed):
        raise ValueError(f"Convolution plan used with a {name} module that had mismatched input/output channels, kernel size, or stride, or transposition")
    x = data.jdata
    relu = norm.activation == "relu"
    training = norm.training or norm.running_mean is None
    fusable = plan.fused_epilogue_available(x, conv.weight) and _norm.native_rows_supported(x.new_empty((1, conv.out_channels)).to(torch.result_type(x, conv.weight)))
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in list(conv.parameters()) + list(norm.parameters())))
    if fusable and not training and not needs_grad:
        with record_function(f"conv_bn_act[inference]({conv!r})"):
            inv = torch.rsqrt(norm.running_var.float() + norm.eps)
            scale = inv * norm.weight.float() if norm.weight is not None else inv
            shift = (norm.bias.float() if norm.bias is not None else 0.0) - norm.running_mean.float() * scale
            return plan.execute_inference(data, conv.weight, conv.bias, scale=scale, shift=shift, residual=residual,
                                          relu=(1 if relu else 0) | (2 if final_relu else 0))
    if fusable and training and x.shape[0] > 0:
        with record_function(f"conv_bn_act[train]({conv!r})"):
            # a bias in front of a training-mode BatchNorm cancels in the normalised output and its gradient is identically
            # zero: the convolution runs without it (no bias add, no column-sum pass over grad_output in backward); only the
            # running mean, which tracks mean(conv(x) + bias), has to see it
            y, stats = plan.execute_with_stats(data, conv.weight, None)
            rows = norm._rows(y.jdata, conv_stats=stats)
            if conv.bias is not None:
                if norm.track_running_stats and norm.running_mean is not None and norm._effective_momentum:
                    with torch.no_grad():
                        norm.running_mean.add_(conv.bias.detach().to(norm.running_mean.dtype), alpha=norm._effective_momentum)
                rows = _ZeroGradFor.apply(rows, conv.bias)
            out = y.jagged_like(rows)
    else:
        out = norm(conv(data, plan))
    if residual is not None:
        out = out.jagged_like(out.jdata + residual.jdata)
    if final_relu:
        out = out.jagged_like(torch.relu(out.jdata))
    return out


class _Pool(nn.Module):
    def __init__(self, kernel_size: NumericMaxRank1, stride: "NumericMaxRank1 | None" = None):
        super().__init__()
        self._kernel_size = to_Vec3i(kernel_size, value_constraint=ValueConstraint.POSITIVE)
        self._stride = to_Vec3i(stride, value_constraint=ValueConstraint.POSITIVE) if stride is not None else self._kernel_size

    kernel_size = property(lambda self: self._kernel_size)
    stride = property(lambda self: self._stride)

    def extra_repr(self) -> str:
        return f"kernel_size={self.kernel_size}, stride={self.stride}"


class MaxPool(_Pool):
    """3-D max pooling of a JaggedTensor over a GridBatch (mirror of reference fvdb/nn/modules.py:114-200)."""

    def forward(self, fine_data: JaggedTensor, fine_grid: GridBatch, coarse_grid: "GridBatch | None" = None):
        with record_function(repr(self)):
            return fine_grid.max_pool(self.kernel_size, fine_data, stride=self.stride, coarse_grid=coarse_grid)


class AvgPool(_Pool):
    """3-D average pooling of a JaggedTensor over a GridBatch (mirror of reference fvdb/nn/modules.py:33-111)."""

    def forward(self, fine_data: JaggedTensor, fine_grid: GridBatch, coarse_grid: "GridBatch | None" = None):
        with record_function(repr(self)):
            return fine_grid.avg_pool(self.kernel_size, fine_data, stride=self.stride, coarse_grid=coarse_grid)


class UpsamplingNearest(nn.Module):
    """Nearest-neighbour upsampling by ``scale_factor`` (mirror of reference fvdb/nn/modules.py:203-260)."""

    def __init__(self, scale_factor: NumericMaxRank1):
        super().__init__()
        self.scale_factor = to_Vec3i(scale_factor, value_constraint=ValueConstraint.POSITIVE)

    def extra_repr(self) -> str:
        return f"scale_factor={self.scale_factor}"

    def forward(self, coarse_data: JaggedTensor, coarse_grid: GridBatch, mask: "JaggedTensor | None" = None, fine_grid: "GridBatch | None" = None):
        with record_function(repr(self)):
            return coarse_grid.refine(self.scale_factor, coarse_data, mask, fine_grid=fine_grid)
