"""Batch normalisation (+ fused ReLU) and column sums over jagged feature rows, on the C ABI of libfvdbconv.so.

The reference applies ``torch.nn.BatchNorm1d`` to ``jdata`` (fvdb/nn/modules.py:484-521) and a separate ReLU pass;
``fvc_bn_*`` (csrc/norm.cu) are streaming kernels for the ``[N, C]`` layout with the activation fused.  Shapes the
kernels do not serve (CPU tensors, fp64, channel counts that are not a multiple of a 16-byte vector) run torch's own
``batch_norm`` -- that *is* the reference implementation of this module.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check, lib

_CODES = {torch.float16: _lib.FVC_F16, torch.bfloat16: _lib.FVC_BF16, torch.float32: _lib.FVC_F32}


def native_rows_supported(x: torch.Tensor) -> bool:
    if not x.is_cuda or x.dim() != 2 or x.dtype not in _CODES or not x.is_contiguous():
        return False
    vec = 4 if x.dtype == torch.float32 else 8
    c = x.shape[1]
    return c > 0 and c % vec == 0 and c // vec <= 256 and x.data_ptr() % 16 == 0


def _stream(x: torch.Tensor) -> int:
    return torch.cuda.current_stream(x.device).cuda_stream


def _scratch(x: torch.Tensor) -> torch.Tensor:
    return torch.empty(int(lib.fvc_bn_scratch_bytes(x.shape[1])), dtype=torch.uint8, device=x.device)


def _f32(t: "torch.Tensor | None") -> "torch.Tensor | None":
    return None if t is None else t.detach().to(torch.float32).contiguous()


def _p(t: "torch.Tensor | None") -> int:
    return 0 if t is None else t.data_ptr()


def column_sums(x: torch.Tensor) -> torch.Tensor:
    """fp32 ``x.sum(0)`` (bias gradient) -- deterministic two-stage reduction."""
    if not native_rows_supported(x):
        return x.float().sum(dim=0)
    sums = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    scratch = _scratch(x)
    with torch.cuda.device(x.device):
        check(lib.fvc_column_sums(x.data_ptr(), x.shape[0], x.shape[1], _CODES[x.dtype], sums.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream(x)))
    return sums


def stats_from_conv_partials(stats, running_mean=None, running_var=None, momentum: float = 0.0):
    """(mean, biased var) over the rows of a convolution output from the per-block column sums its fused epilogue wrote
    (``_fvdb_cpp.ConvStats``): the BatchNorm statistics pass without reading ``y`` again.  Optionally updates fp32 running
    statistics in place like ``fvc_bn_stats``."""
    partial = stats.partial
    c = int(partial.shape[2])
    mean = torch.empty(c, dtype=torch.float32, device=partial.device)
    var = torch.empty(c, dtype=torch.float32, device=partial.device)
    blocks = (stats.rows + stats.rows_per_block - 1) // stats.rows_per_block
    with torch.cuda.device(partial.device):
        check(lib.fvc_bn_stats_from_partials(partial.data_ptr(), blocks, stats.rows_per_block, stats.rows, c, mean.data_ptr(), var.data_ptr(),
                                             _p(running_mean), _p(running_var), float(momentum), torch.cuda.current_stream(partial.device).cuda_stream))
    return mean, var


class BatchNormFn(torch.autograd.Function):
    """y = act(batch_norm(x)); statistics over the rows of this process, or of every process of ``group`` (SyncBatchNorm)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, relu, group, conv_stats=None):  # type: ignore[override]
        """``conv_stats``: the column sums the producing convolution's epilogue wrote for ``x`` (``_fvdb_cpp.ConvStats``); the
        statistics pass over ``x`` is skipped."""
        n, c = x.shape
        code = _CODES[x.dtype]
        gamma, beta = _f32(weight), _f32(bias)
        scratch = _scratch(x)
        count, count_dev = n, None
        with torch.cuda.device(x.device):
            stream = _stream(x)
            if training:
                mean = torch.empty(c, dtype=torch.float32, device=x.device)
                var = torch.empty(c, dtype=torch.float32, device=x.device)
                in_place = group is None and running_mean is not None and running_mean.dtype == torch.float32 and running_var.dtype == torch.float32
                if conv_stats is not None and n > 0:
                    blocks = (conv_stats.rows + conv_stats.rows_per_block - 1) // conv_stats.rows_per_block
                    check(lib.fvc_bn_stats_from_partials(conv_stats.partial.data_ptr(), blocks, conv_stats.rows_per_block, n, c, mean.data_ptr(), var.data_ptr(),
                                                         _p(running_mean) if in_place else 0, _p(running_var) if in_place else 0, float(momentum), stream))
                else:
                    check(lib.fvc_bn_stats(x.data_ptr(), n, c, code, mean.data_ptr(), var.data_ptr(), _p(running_mean) if in_place else 0,
                                           _p(running_var) if in_place else 0, float(momentum), scratch.data_ptr(), scratch.numel(), stream))
                if group is not None:  # merge (count, mean, M2) of every rank
                    import torch.distributed as dist

                    packed = torch.cat([mean, var * n, torch.full((1,), float(n), device=x.device)])
                    gathered = [torch.empty_like(packed) for _ in range(dist.get_world_size(group))]
                    dist.all_gather(gathered, packed, group=group)
                    g = torch.stack(gathered).double()
                    counts = g[:, -1:]
                    total = counts.sum().clamp_min(1)  # (an all-empty batch normalises nothing)
                    mean64 = (g[:, :c] * counts).sum(0) / total
                    m2 = g[:, c:2 * c].sum(0) + (counts * (g[:, :c] - mean64) ** 2).sum(0)
                    mean, var, count_dev = mean64.float(), (m2 / total).float(), total.float().reshape(1)  # the count stays on the device
                if running_mean is not None and not in_place:
                    unbiased = var * (count_dev / (count_dev - 1).clamp_min(1)) if count_dev is not None else var * (count / max(count - 1, 1))
                    running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                    running_var.mul_(1 - momentum).add_(unbiased.to(running_var.dtype), alpha=momentum)
            else:
                mean, var = _f32(running_mean), _f32(running_var)
            y = torch.empty_like(x)
            check(lib.fvc_bn_apply(x.data_ptr(), n, c, code, mean.data_ptr(), var.data_ptr(), _p(gamma), _p(beta), float(eps), int(relu), y.data_ptr(), stream))
        ctx.save_for_backward(x, mean, var, gamma if gamma is not None else x.new_empty(0), beta if beta is not None else x.new_empty(0),
                              count_dev if count_dev is not None else x.new_empty(0))
        ctx.cfg = (bool(training), float(eps), bool(relu), group, count, weight is not None, bias is not None,
                   None if weight is None else weight.dtype, None if bias is None else bias.dtype)
        return y

    @staticmethod
    def backward(ctx, grad_output):  # type: ignore[override]
        x, mean, var, gamma, beta, count_dev = ctx.saved_tensors
        training, eps, relu, group, count, has_w, has_b, w_dtype, b_dtype = ctx.cfg
        gamma = gamma if has_w else None
        beta = beta if has_b else None
        n, c = x.shape
        code = _CODES[x.dtype]
        dy = grad_output.contiguous()
        sums = torch.empty((2, c), dtype=torch.float32, device=x.device)
        scratch = _scratch(x)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(x.device):
            stream = _stream(x)
            check(lib.fvc_bn_backward_reduce(dy.data_ptr(), x.data_ptr(), n, c, code, mean.data_ptr(), var.data_ptr(), _p(gamma), _p(beta), eps, int(relu),
                                             sums.data_ptr(), scratch.data_ptr(), scratch.numel(), stream))
            local = sums
            if group is not None and training:
                import torch.distributed as dist

                sums = sums.clone()
                dist.all_reduce(sums, group=group)
            if dx is not None:
                check(lib.fvc_bn_backward_apply(dy.data_ptr(), x.data_ptr(), n, c, code, mean.data_ptr(), var.data_ptr(), _p(gamma), _p(beta), eps, int(relu),
                                                int(training), sums.data_ptr(), count, count_dev.data_ptr() if count_dev.numel() else 0, dx.data_ptr(), stream))
        grad_w = local[1].to(w_dtype) if has_w and ctx.needs_input_grad[1] else None
        grad_b = local[0].to(b_dtype) if has_b and ctx.needs_input_grad[2] else None
        return dx, grad_w, grad_b, None, None, None, None, None, None, None, None


class _SyncBatchNormTorchFn(torch.autograd.Function):
    """Distributed batch norm in plain torch ops for the shapes the native kernels do not serve (fp64, odd channel counts):
    same collectives as the native path (one all_gather of (mean, M2, count) forward, one all_reduce of the two sums
    backward), so every rank of the group -- including a rank with zero rows -- takes part."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu, group):  # type: ignore[override]
        import torch.distributed as dist

        n, c = x.shape
        xd = x.double()
        mean_l = xd.mean(0) if n else xd.new_zeros(c)
        m2_l = ((xd - mean_l) ** 2).sum(0) if n else xd.new_zeros(c)
        packed = torch.cat([mean_l, m2_l, xd.new_full((1,), float(n))])
        gathered = [torch.empty_like(packed) for _ in range(dist.get_world_size(group))]
        dist.all_gather(gathered, packed, group=group)
        g = torch.stack(gathered)
        counts = g[:, -1:]
        total = counts.sum().clamp_min(1)
        mean = (g[:, :c] * counts).sum(0) / total
        var = (g[:, c:2 * c].sum(0) + (counts * (g[:, :c] - mean) ** 2).sum(0)) / total
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
            running_var.mul_(1 - momentum).add_((var * total / (total - 1).clamp_min(1)).to(running_var.dtype), alpha=momentum)
        invstd = torch.rsqrt(var + eps)
        xhat = (xd - mean) * invstd
        y = xhat * (weight.double() if weight is not None else 1.0) + (bias.double() if bias is not None else 0.0)
        if relu:
            y = torch.relu(y)
        ctx.save_for_backward(xhat, invstd, weight if weight is not None else x.new_empty(0), (y > 0) if relu else x.new_empty(0), total)
        ctx.cfg = (group, relu, weight is not None, bias is not None, x.dtype)
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, grad_output):  # type: ignore[override]
        import torch.distributed as dist

        xhat, invstd, weight, positive, total = ctx.saved_tensors
        group, relu, has_w, has_b, dtype = ctx.cfg
        dz = grad_output.double()
        if relu:
            dz = dz * positive
        local = torch.stack([dz.sum(0), (dz * xhat).sum(0)])
        sums = local.clone()
        dist.all_reduce(sums, group=group)
        gamma = weight.double() if has_w else 1.0
        dx = (gamma * invstd * (dz - sums[0] / total - xhat * sums[1] / total)).to(dtype)
        return dx, (local[1].to(weight.dtype) if has_w else None), (local[0].to(dtype) if has_b else None), None, None, None, None, None, None


def batch_norm_rows(x, weight, bias, running_mean, running_var, training, momentum, eps, relu=False, group=None, conv_stats=None):
    """Functional form used by fvdb.nn.BatchNorm / SyncBatchNorm.

    Which implementation runs depends only on rank-invariant properties (dtype, channel count, device): with a process
    group every rank must reach the same collectives, also a rank that owns zero rows (the by-grid partition leaves ranks
    empty when there are fewer grids than ranks; torch.nn.SyncBatchNorm, which the reference subclasses, accepts that)."""
    distributed = group is not None and training
    if native_rows_supported(x) and (training or running_mean is not None) and (x.shape[0] > 0 or distributed):
        return BatchNormFn.apply(x, weight, bias, running_mean, running_var, training, momentum, eps, relu, group, conv_stats if training else None)
    if distributed:
        return _SyncBatchNormTorchFn.apply(x.contiguous(), weight, bias, running_mean, running_var, momentum, eps, relu, group)
    y = F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)
    return torch.relu(y) if relu else y
