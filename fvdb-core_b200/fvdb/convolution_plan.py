"""``ConvolutionPlan``: a reusable, immutable description of one sparse 3-D convolution over a ``GridBatch``.

API-for-API mirror of reference fvdb/convolution_plan.py (factories :640-908, ``valid_usage`` :914, ``execute`` :961,
properties :1086-1161, backend choice :1167-1219) so that code written against fVDB runs unchanged; the
execution underneath is this repo's B200 engine (``fvdb._fvdb_cpp`` -> libfvdbconv.so).  A plan stores the
finite relation ``fine = stride * coarse + tap - floor((kernel_size - 1) / 2)`` restricted to its two grids.
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass
from threading import Lock
from typing import Any

import torch

from . import _fvdb_cpp, _norm
from .enums import ConvolutionPhasePolicy, ConvolutionTopologyPolicy, ConvolutionTopologyProvenance
from .grid_batch import GridBatch
from .jagged_tensor import JaggedTensor
from .types import NumericMaxRank1, ValueConstraint, to_Vec3i

_DEFAULT_CONFIG: dict[str, Any] = {"backend": "default"}
_ANY_CHANNEL_PAIRS: tuple[tuple[int, int], ...] = ()
_TRANSFORM_COMPATIBILITY_ATOL = 1.0e-6
_TRANSFORM_COMPATIBILITY_RTOL = 1.0e-6
_WARNED_INCOMPLETE_COVERAGE_GEOMETRIES: set[tuple[tuple[int, int, int], tuple[int, int, int]]] = set()
_DENSE_BACKEND_DISABLED_MESSAGE = (
    "The dense convolution backend is disabled because it does not yet implement the canonical sparse geometry "
    "and public transposed-weight layout; use backend='default' or backend='gather_scatter'."
)
_MIGRATION_HINT = (
    "Rebuild the explicit target with stride-scaled voxel sizes and identical origins. "
    "GridBatch.coarsened_grid uses a different block-centroid transform contract."
)


class ConvolutionCoverageWarning(UserWarning):
    """Some stride residues of this (kernel_size, stride) geometry are never sampled; emitted once per geometry."""


@dataclass(frozen=True)
class ConvolutionCoverageReport:
    """Exact rulebook degree statistics of a plan (rows = voxels, degree = number of edges)."""

    input_row_count: int
    output_row_count: int
    input_zero_count: int
    input_zero_fraction: float
    input_degree_min: int
    input_degree_max: int
    input_degree_histogram: tuple[tuple[int, int], ...]
    output_zero_count: int
    output_zero_fraction: float
    output_degree_min: int
    output_degree_max: int
    output_degree_histogram: tuple[tuple[int, int], ...]


@dataclass(frozen=True)
class ConvolutionTransformCompatibility:
    """Whether the fine and coarse lattices satisfy ``h_coarse == stride * h_fine`` with registration ``a == 0``."""

    fine_grid_count: int
    coarse_grid_count: int
    same_batch_size: bool
    same_device: bool
    scale_compatible: bool
    registration_integer: bool
    registration_zero: bool
    compatible: bool
    registration_offset: torch.Tensor | None


def _close(a: torch.Tensor, b: torch.Tensor) -> bool:
    return bool(torch.allclose(a, b, atol=_TRANSFORM_COMPATIBILITY_ATOL, rtol=_TRANSFORM_COMPATIBILITY_RTOL))


def _transform_compatibility(fine_grid: GridBatch, coarse_grid: GridBatch, geometry) -> ConvolutionTransformCompatibility:
    same_batch = fine_grid.grid_count == coarse_grid.grid_count
    same_device = fine_grid.device == coarse_grid.device
    if not (same_batch and same_device):
        return ConvolutionTransformCompatibility(
            fine_grid.grid_count, coarse_grid.grid_count, same_batch, same_device, False, False, False, False, None
        )
    # float64 host metadata is authoritative (the float32 GridBatch view would lose registration precision)
    h_fine, h_coarse = fine_grid.data.voxel_sizes, coarse_grid.data.voxel_sizes
    scale_ok = _close(h_coarse, h_fine * torch.tensor(geometry.stride, dtype=h_fine.dtype))
    offset = (coarse_grid.data.origins - fine_grid.data.origins) / h_fine
    is_integer = _close(offset, torch.round(offset))
    is_zero = _close(offset, torch.zeros_like(offset))
    return ConvolutionTransformCompatibility(
        fine_grid.grid_count, coarse_grid.grid_count, True, True, scale_ok, is_integer, is_zero, scale_ok and is_integer and is_zero, offset
    )


def _validate_transform_compatibility(c: ConvolutionTransformCompatibility) -> None:
    if not c.same_batch_size:
        raise ValueError(
            f"Convolution fine and coarse grids must have the same batch size; got {c.fine_grid_count} and {c.coarse_grid_count}. {_MIGRATION_HINT}"
        )
    if not c.same_device:
        raise ValueError(f"Convolution fine and coarse grids must be on the same device. {_MIGRATION_HINT}")
    if not c.scale_compatible:
        raise ValueError(f"Convolution voxel size mismatch: expected h_coarse = stride * h_fine in every batch and axis. {_MIGRATION_HINT}")
    if c.registration_offset is None or not bool(torch.isfinite(c.registration_offset).all()):
        raise ValueError(f"Convolution registration offset must be finite. {_MIGRATION_HINT}")
    if not c.registration_integer:
        raise ValueError(f"Convolution grids have a fractional lattice registration offset. Convolution currently supports only a=0. {_MIGRATION_HINT}")
    if not c.registration_zero:
        raise ValueError(f"Convolution grids have a nonzero integer lattice registration offset. Convolution currently supports only a=0. {_MIGRATION_HINT}")


def _resolve_topology_policy(target_grid, topology_policy) -> ConvolutionTopologyPolicy:
    if topology_policy is None:
        return ConvolutionTopologyPolicy.COMPLETE if target_grid is None else ConvolutionTopologyPolicy.RESTRICTED
    if not isinstance(topology_policy, ConvolutionTopologyPolicy):
        raise TypeError("topology_policy must be a ConvolutionTopologyPolicy value")
    if topology_policy is ConvolutionTopologyPolicy.COMPLETE and target_grid is not None:
        raise ValueError("topology_policy=ConvolutionTopologyPolicy.COMPLETE requires target_grid=None")
    if topology_policy is ConvolutionTopologyPolicy.RESTRICTED and target_grid is None:
        raise ValueError("topology_policy=ConvolutionTopologyPolicy.RESTRICTED requires an explicit target_grid")
    return topology_policy


def _warn_if_incomplete_residue_coverage(geometry, acknowledged: bool) -> None:
    if acknowledged:
        return
    ks, st, pad = tuple(geometry.kernel_size), tuple(geometry.stride), geometry.padding_before
    uncovered = [axis for axis in range(3) if len({(tap - pad[axis]) % st[axis] for tap in range(ks[axis])}) != st[axis]]
    key = (ks, st)
    if not uncovered or key in _WARNED_INCOMPLETE_COVERAGE_GEOMETRIES:
        return
    _WARNED_INCOMPLETE_COVERAGE_GEOMETRIES.add(key)
    warnings.warn(
        f"This convolution geometry leaves uncovered stride residues on axes {uncovered}; some active fine coordinates can have "
        "zero rulebook degree. This matches dense Torch sampling. Pass acknowledge_incomplete_coverage=True to suppress this warning.",
        ConvolutionCoverageWarning,
        stacklevel=3,
    )


def _degree_summary(degrees: torch.Tensor):
    rows = int(degrees.numel())
    if rows == 0:
        return 0, 0.0, 0, 0, ()
    degrees = degrees.cpu()
    zeros = int((degrees == 0).sum())
    values, counts = torch.unique(degrees, sorted=True, return_counts=True)
    return zeros, zeros / rows, int(degrees.min()), int(degrees.max()), tuple(zip(values.tolist(), counts.tolist()))


# ---- backends (cached per-plan precomputed data) ----------------------------------------------------


@dataclass(frozen=True)
class _MatmulBackend:
    """K = S = 1 on one shared grid: plain matmul, no kernel map."""


@dataclass(frozen=True)
class _GatherScatterBackend:
    """Kernel-map convolution (the B200 engine)."""

    topology: _fvdb_cpp.GatherScatterDefaultTopology


@dataclass(frozen=True)
class _PredGatherIGemmBackend:
    """Reference name for its tensor-core forward backend; here the same engine, kept for admission parity."""

    gs_topology: _fvdb_cpp.GatherScatterDefaultTopology
    kernel_size: int
    stride: int


_Backend = _MatmulBackend | _GatherScatterBackend | _PredGatherIGemmBackend


def _backend_topology(backend: _Backend):
    if isinstance(backend, _GatherScatterBackend):
        return backend.topology
    if isinstance(backend, _PredGatherIGemmBackend):
        return backend.gs_topology
    return None


def _coverage_report(backend: _Backend, source_grid: GridBatch, target_grid: GridBatch) -> ConvolutionCoverageReport | None:
    if isinstance(backend, _MatmulBackend):
        n_in, n_out = source_grid.total_voxels, target_grid.total_voxels
        one_in, one_out = (1 if n_in else 0), (1 if n_out else 0)
        return ConvolutionCoverageReport(
            n_in, n_out, 0, 0.0, one_in, one_in, ((1, n_in),) if n_in else (), 0, 0.0, one_out, one_out, ((1, n_out),) if n_out else ()
        )
    topology = _backend_topology(backend)
    if topology is None:
        return None
    d_in = torch.bincount(topology.gather_indices, minlength=topology.feature_total_voxels)
    d_out = torch.bincount(topology.scatter_indices, minlength=topology.output_total_voxels)
    return ConvolutionCoverageReport(int(d_in.numel()), int(d_out.numel()), *_degree_summary(d_in), *_degree_summary(d_out))


def _swap_coverage_report(r: ConvolutionCoverageReport) -> ConvolutionCoverageReport:
    return ConvolutionCoverageReport(
        r.output_row_count, r.input_row_count,
        r.output_zero_count, r.output_zero_fraction, r.output_degree_min, r.output_degree_max, r.output_degree_histogram,
        r.input_zero_count, r.input_zero_fraction, r.input_degree_min, r.input_degree_max, r.input_degree_histogram,
    )


def _output_zero_count(backend: _Backend) -> int | None:
    if isinstance(backend, _MatmulBackend):
        return 0
    topology = _backend_topology(backend)
    if topology is None:
        return None
    view = topology._core
    if not topology._reversed:
        return view.zero_degree_outputs()  # from the dense map: the CSR arrays are not materialised for a policy check
    return int((torch.bincount(topology.scatter_indices, minlength=topology.output_total_voxels) == 0).sum())


def _validate_coverage_policy(backend: _Backend, policy: ConvolutionTopologyPolicy, strict: bool) -> None:
    if policy is not ConvolutionTopologyPolicy.COMPLETE and not strict:
        return
    zeros = _output_zero_count(backend)
    if not zeros:
        return
    if policy is ConvolutionTopologyPolicy.COMPLETE:
        raise RuntimeError(f"Generated complete topology contains {zeros} zero-degree output rows.")
    if strict:
        raise ValueError(f"Restricted topology contains {zeros} zero-degree output rows.")


def _channel_pair_supported(cin: int, cout: int, pairs) -> bool:
    return len(pairs) == 0 or (cin, cout) in pairs


def _pred_gather_igemm_channel_pair_supported(cin: int, cout: int) -> bool:
    return cin > 0 and cout > 0 and cin % 32 == 0 and cout % 32 == 0


def _validate_pred_gather_igemm_admission(kernel_size, stride, channel_pairs, *, transposed: bool) -> tuple[int, int]:
    """The reference's admission boundary for backend='pred_gather_igemm' (convolution_plan.py:376-404)."""
    if transposed:
        raise ValueError("PredGatherIGemm backend does not support transposed convolution.")
    ks = [int(v) for v in kernel_size.tolist()]
    if len(set(ks)) != 1 or ks[0] not in (3, 5, 7):
        raise ValueError(f"PredGatherIGemm supports only uniform kernel sizes 3, 5, 7; got {ks}.")
    st = [int(v) for v in stride.tolist()]
    if len(set(st)) != 1 or st[0] not in (1, 2):
        raise ValueError(f"PredGatherIGemm supports only uniform strides 1, 2; got {st}.")
    for pair in channel_pairs:
        if len(pair) != 2 or pair[0] <= 0 or pair[1] <= 0:
            raise ValueError("channel_pair must be a tuple of two positive integers")
        if not _pred_gather_igemm_channel_pair_supported(*pair):
            raise ValueError(f"PredGatherIGemm requires channel counts divisible by 32; got ({pair[0]}, {pair[1]}).")
    return ks[0], st[0]


def _validate_pred_gather_igemm_grid_admission(source_grid: GridBatch, target_grid: GridBatch | None = None) -> None:
    for grid, which in ((source_grid, "source"), (target_grid, "target")):
        if grid is None:
            continue
        if grid.device.type != "cuda":
            raise ValueError("PredGatherIGemm requires source and target grids on CUDA.")
        if grid.grid_count != 1:
            raise ValueError(f"PredGatherIGemm supports only batch size 1; got {grid.grid_count} {which} grids.")


def _matmul_weight_matrix(weights: torch.Tensor) -> torch.Tensor:
    if weights.ndim == 2:
        return weights
    if weights.ndim == 5 and tuple(weights.shape[2:]) == (1, 1, 1):
        return weights[:, :, 0, 0, 0]
    raise ValueError("The K=1, S=1 matmul backend requires weights shaped [C_out, C_in] or [C_out, C_in, 1, 1, 1].")


class _GatherScatterConvFn(torch.autograd.Function):
    """autograd glue over gs_conv / gs_conv_backward (either direction); an optional bias is added in the kernel epilogue."""

    @staticmethod
    def forward(ctx, features, weights, bias, topo, transposed, stats_out=None):  # type: ignore[override]
        """``stats_out`` (a dict, extension): the kernel epilogue also writes per-block column sums of the stored output; the
        ``_fvdb_cpp.ConvStats`` lands in ``stats_out["stats"]`` (not differentiable: it feeds BatchNorm statistics)."""
        fn = _fvdb_cpp.gs_conv_transpose if transposed else _fvdb_cpp.gs_conv
        # fp32 on the tensor pipe: the layer's input is split into its three bf16 parts ONCE; the split rows serve the
        # forward pass now and the weight gradient later (instead of a 10 B/element pre-pass in each of the two calls)
        split = None
        if (features.dtype == torch.float32 and weights.dtype == torch.float32 and features.is_cuda and features.shape[0] > 0
                and (ctx.needs_input_grad[1] or ctx.needs_input_grad[0])
                and _fvdb_cpp._tensor_core_fp32(int(weights.shape[1]), int(weights.shape[0]), topo.kernel_volume)):
            with torch.cuda.device(features.device):
                split = _fvdb_cpp.split_rows(features.contiguous())
        ctx.save_for_backward(features, weights, *([split] if split is not None else []))
        ctx.topo, ctx.transposed, ctx.has_bias = topo, transposed, bias is not None
        if stats_out is not None:
            y, stats = fn(features, weights, topo, bias, features_split=split, want_stats=True)
            stats_out["stats"] = stats
            return y
        return fn(features, weights, topo, bias, features_split=split)

    @staticmethod
    def backward(ctx, grad_output):  # type: ignore[override]
        features, weights, *rest = ctx.saved_tensors
        fn = _fvdb_cpp.gs_conv_transpose_backward if ctx.transposed else _fvdb_cpp.gs_conv_backward
        grad_output = grad_output.contiguous()
        grad_features, grad_weights = fn(grad_output, features, weights, ctx.topo, features_split=rest[0] if rest else None,
                                         need_grad_features=ctx.needs_input_grad[0])
        # bias gradient: fp32 column sums in one streaming pass (csrc/norm.cu), rounded once
        grad_bias = _norm.column_sums(grad_output).to(grad_output.dtype) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return grad_features, grad_weights, grad_bias, None, None, None


class _CoverageReportCache:
    """Lazy, thread-safe coverage diagnostics shared between a plan and its exact transposes."""

    def __init__(self, backend: _Backend, source_grid: GridBatch, target_grid: GridBatch):
        self._args = (backend, source_grid, target_grid)
        self._report: ConvolutionCoverageReport | None = None
        self._swapped: ConvolutionCoverageReport | None = None
        self._lock = Lock()

    def get(self, swapped: bool) -> ConvolutionCoverageReport | None:
        with self._lock:
            if self._report is None:
                self._report = _coverage_report(*self._args)
            if not swapped or self._report is None:
                return self._report
            if self._swapped is None:
                self._swapped = _swap_coverage_report(self._report)
            return self._swapped


@dataclass(frozen=True)
class ConvolutionPlan:
    """Create with ``from_grid_batch`` / ``from_grid_batch_transposed`` / ``from_plan_transposed``; run with ``execute``."""

    _source_grid: GridBatch
    _target_grid: GridBatch
    _geometry: _fvdb_cpp.ConvolutionGeometry
    _channel_pairs: tuple[tuple[int, int], ...]
    _transposed: bool
    _backend: _Backend
    _transform_compatibility: ConvolutionTransformCompatibility
    _topology_policy: ConvolutionTopologyPolicy
    _topology_provenance: ConvolutionTopologyProvenance
    _coverage_report_cache: _CoverageReportCache
    _coverage_report_swapped: bool

    # ---- factories ------------------------------------------------------------------------
    @classmethod
    def _from_grids(cls, transposed, kernel_size, stride, source_grid, target_grid, expert_config, channel_pairs, topology_policy,
                    strict_output_coverage, acknowledge_incomplete_coverage) -> "ConvolutionPlan":
        kernel_size = to_Vec3i(kernel_size, value_constraint=ValueConstraint.POSITIVE)
        stride = to_Vec3i(stride, value_constraint=ValueConstraint.POSITIVE)
        policy = _resolve_topology_policy(target_grid, topology_policy)
        provenance = (
            ConvolutionTopologyProvenance.GENERATED if policy is ConvolutionTopologyPolicy.COMPLETE else ConvolutionTopologyProvenance.EXPLICIT_TARGET
        )
        backend_name = expert_config.get("backend", "default")
        if backend_name == "dense":
            raise ValueError(_DENSE_BACKEND_DISABLED_MESSAGE)
        if backend_name == "pred_gather_igemm":
            _validate_pred_gather_igemm_admission(kernel_size, stride, channel_pairs, transposed=transposed)
            _validate_pred_gather_igemm_grid_admission(source_grid, target_grid)
        if target_grid is None:
            target_grid = source_grid.conv_transpose_grid(kernel_size, stride) if transposed else source_grid.conv_grid(kernel_size, stride)
        geometry = _fvdb_cpp.ConvolutionGeometry(kernel_size, stride)
        fine, coarse = (target_grid, source_grid) if transposed else (source_grid, target_grid)
        compatibility = _transform_compatibility(fine, coarse, geometry)
        _validate_transform_compatibility(compatibility)
        _warn_if_incomplete_residue_coverage(geometry, acknowledge_incomplete_coverage)
        backend = cls._build_backend(source_grid, target_grid, kernel_size, stride, channel_pairs, expert_config, transposed=transposed)
        _validate_coverage_policy(backend, policy, strict_output_coverage)
        return cls(source_grid, target_grid, geometry, channel_pairs, transposed, backend, compatibility, policy, provenance,
                   _CoverageReportCache(backend, source_grid, target_grid), False)

    @classmethod
    def from_grid_batch(cls, kernel_size: NumericMaxRank1, stride: NumericMaxRank1, source_grid: GridBatch, target_grid: GridBatch | None = None, *,
                        expert_config: dict[str, Any] = _DEFAULT_CONFIG, channel_pairs: tuple[tuple[int, int], ...] = _ANY_CHANNEL_PAIRS,
                        topology_policy: ConvolutionTopologyPolicy | None = None, strict_output_coverage: bool = False,
                        acknowledge_incomplete_coverage: bool = False) -> "ConvolutionPlan":
        """Plan a convolution from ``source_grid`` to ``target_grid`` (generated with ``conv_grid`` when ``None``)."""
        return cls._from_grids(False, kernel_size, stride, source_grid, target_grid, expert_config, channel_pairs, topology_policy,
                               strict_output_coverage, acknowledge_incomplete_coverage)

    @classmethod
    def from_grid_batch_transposed(cls, kernel_size: NumericMaxRank1, stride: NumericMaxRank1, source_grid: GridBatch, target_grid: GridBatch | None = None, *,
                                   expert_config: dict[str, Any] = _DEFAULT_CONFIG, channel_pairs: tuple[tuple[int, int], ...] = _ANY_CHANNEL_PAIRS,
                                   topology_policy: ConvolutionTopologyPolicy | None = None, strict_output_coverage: bool = False,
                                   acknowledge_incomplete_coverage: bool = False) -> "ConvolutionPlan":
        """Plan a transposed convolution (target generated with ``conv_transpose_grid`` when ``None``)."""
        return cls._from_grids(True, kernel_size, stride, source_grid, target_grid, expert_config, channel_pairs, topology_policy,
                               strict_output_coverage, acknowledge_incomplete_coverage)

    @classmethod
    def from_plan_transposed(cls, plan: "ConvolutionPlan") -> "ConvolutionPlan":
        """The exact finite adjoint connectivity of ``plan``: no device work, index tensors aliased."""
        source_grid, target_grid, transposed = plan._target_grid, plan._source_grid, not plan._transposed
        fine, coarse = (target_grid, source_grid) if transposed else (source_grid, target_grid)
        compatibility = _transform_compatibility(fine, coarse, plan._geometry)
        _validate_transform_compatibility(compatibility)
        if isinstance(plan._backend, _MatmulBackend):
            backend: _Backend = plan._backend
        elif isinstance(plan._backend, (_GatherScatterBackend, _PredGatherIGemmBackend)):
            backend = _GatherScatterBackend(topology=_fvdb_cpp.gs_reverse_topology(_backend_topology(plan._backend)))
        else:
            raise TypeError(f"Cannot transpose unknown convolution backend: {type(plan._backend)}")
        return cls(source_grid, target_grid, plan._geometry, tuple((dst, src) for src, dst in plan._channel_pairs), transposed, backend,
                   compatibility, ConvolutionTopologyPolicy.RESTRICTED, ConvolutionTopologyProvenance.EXACT_TRANSPOSE,
                   plan._coverage_report_cache, not plan._coverage_report_swapped)

    # ---- validation -----------------------------------------------------------------------
    def valid_usage(self, in_channels: int, out_channels: int, kernel_size: NumericMaxRank1, stride: NumericMaxRank1, transposed: bool) -> bool:
        kernel_size = to_Vec3i(kernel_size, value_constraint=ValueConstraint.POSITIVE)
        stride = to_Vec3i(stride, value_constraint=ValueConstraint.POSITIVE)
        backend_ok = not isinstance(self._backend, _PredGatherIGemmBackend) or _pred_gather_igemm_channel_pair_supported(in_channels, out_channels)
        return (
            _channel_pair_supported(in_channels, out_channels, self._channel_pairs)
            and backend_ok
            and kernel_size.tolist() == self._geometry.kernel_size
            and stride.tolist() == self._geometry.stride
            and transposed == self._transposed
        )

    # ---- execution ------------------------------------------------------------------------
    def execute(self, data: JaggedTensor | torch.Tensor, weights: torch.Tensor, bias: torch.Tensor | None = None) -> JaggedTensor | torch.Tensor:
        """Apply the planned convolution.  ``data`` is a JaggedTensor (or a plain ``[N, Cin]`` tensor when the
        batch holds one grid); the result has the same kind.  ``weights``: ``[Cout, Cin, k0, k1, k2]``
        (``[Cout, Cin]`` also accepted by K=S=1 plans).  ``bias`` (extension): fused into the kernel epilogue."""
        assert isinstance(data, (torch.Tensor, JaggedTensor)), "data must be a torch.Tensor or JaggedTensor"
        assert isinstance(weights, torch.Tensor), "weights must be a torch.Tensor"
        backend = self._backend
        if isinstance(backend, _MatmulBackend):
            weight_matrix = _matmul_weight_matrix(weights)
            out_c, in_c = weight_matrix.shape
        else:
            if weights.ndim < 2:
                raise ValueError("Convolution weights must have output and input channel dimensions")
            out_c, in_c = weights.shape[0], weights.shape[1]
        if not _channel_pair_supported(in_c, out_c, self._channel_pairs):
            raise ValueError(f"Channel pair {in_c, out_c} is not supported")
        if isinstance(backend, _PredGatherIGemmBackend) and not _pred_gather_igemm_channel_pair_supported(in_c, out_c):
            raise ValueError(f"PredGatherIGemm requires input and output channel counts divisible by 32; got ({in_c}, {out_c}).")
        is_flat = isinstance(data, torch.Tensor)
        if is_flat and self._source_grid.grid_count != 1:
            raise ValueError("Source grid must have batch size of 1 for flat data")

        if isinstance(backend, _MatmulBackend):
            features = data if is_flat else data.jdata
            out = features.matmul(weight_matrix.transpose(0, 1))
            if bias is not None:
                out = out + bias
            return out if is_flat else data.jagged_like(out)

        features = data if is_flat else data.jdata
        topology = _backend_topology(backend)
        if topology is None:
            raise TypeError(f"Unknown backend type: {type(backend)}")
        out = _GatherScatterConvFn.apply(features, weights, bias, topology, self._transposed)
        return out if is_flat else self._target_grid.jagged_like(out)

    def fused_epilogue_available(self, features: torch.Tensor, weights: torch.Tensor) -> bool:
        """Whether ``execute_with_stats`` / ``execute_inference`` can run for these operands: a kernel-map plan on CUDA whose
        (dtype, channels, kernel volume) the tensor-core executor admits (the fused block epilogue lives there)."""
        topology = _backend_topology(self._backend)
        if topology is None or isinstance(self._backend, _MatmulBackend) or not features.is_cuda or weights.ndim != 5:
            return False
        working = torch.result_type(features, weights)
        if working not in _fvdb_cpp._DTYPE_CODE:
            return False
        return int(_fvdb_cpp.lib.fvc_conv_kernel_family(int(weights.shape[1]), int(weights.shape[0]), topology.kernel_volume, _fvdb_cpp._DTYPE_CODE[working], _fvdb_cpp._path, 0)) == 2

    def execute_with_stats(self, data: JaggedTensor | torch.Tensor, weights: torch.Tensor, bias: torch.Tensor | None = None):
        """``execute`` (differentiable) that also returns the per-block column sums of the output written by the kernel
        epilogue (``_fvdb_cpp.ConvStats``): a following BatchNorm takes its batch statistics from them instead of reading
        the output again (reference composition fvdb/nn/modules.py:484-521)."""
        is_flat = isinstance(data, torch.Tensor)
        features = data if is_flat else data.jdata
        holder: dict = {}
        out = _GatherScatterConvFn.apply(features, weights, bias, _backend_topology(self._backend), self._transposed, holder)
        return (out if is_flat else self._target_grid.jagged_like(out)), holder["stats"]

    def execute_inference(self, data: JaggedTensor | torch.Tensor, weights: torch.Tensor, bias: torch.Tensor | None = None, *, scale=None, shift=None,
                          residual=None, relu: "bool | int" = False):
        """Forward only (no autograd): ``act2(act1((conv(x) + bias) * scale + shift) + residual)`` in ONE kernel -- an eval-mode
        BatchNorm folds into ``scale`` / ``shift`` (fp32 ``[Cout]``), the block's ReLU (``relu`` bit 0: act1), the skip connection
        and the ReLU after it (bit 1: act2; fvdb/nn/simple_unet.py:187-188) into the epilogue."""
        is_flat = isinstance(data, torch.Tensor)
        features = (data if is_flat else data.jdata).detach()
        res = None if residual is None else (residual if isinstance(residual, torch.Tensor) else residual.jdata).detach()
        fn = _fvdb_cpp.gs_conv_transpose if self._transposed else _fvdb_cpp.gs_conv
        with torch.no_grad():
            out = fn(features, weights.detach(), _backend_topology(self._backend), None if bias is None else bias.detach(), scale=scale, shift=shift,
                     residual=res, relu=relu)
        return out if is_flat else self._target_grid.jagged_like(out)

    # ---- properties -----------------------------------------------------------------------
    source_grid_batch = property(lambda self: self._source_grid)
    target_grid_batch = property(lambda self: self._target_grid)
    geometry = property(lambda self: self._geometry)
    kernel_size = property(lambda self: torch.tensor(self._geometry.kernel_size, dtype=torch.int32))
    stride = property(lambda self: torch.tensor(self._geometry.stride, dtype=torch.int32))
    transform_compatibility = property(lambda self: self._transform_compatibility)
    phase_policy = property(lambda self: ConvolutionPhasePolicy(self._geometry.phase_policy))
    topology_policy = property(lambda self: self._topology_policy)
    topology_provenance = property(lambda self: self._topology_provenance)
    coverage_report = property(lambda self: self._coverage_report_cache.get(self._coverage_report_swapped))
    has_fixed_topology = property(lambda self: self._source_grid.data.is_same(self._target_grid.data))

    # ---- backend choice -------------------------------------------------------------------
    @staticmethod
    def _build_backend(source_grid, target_grid, kernel_size, stride, channel_pairs, expert_config, transposed: bool = False) -> _Backend:
        backend_name = expert_config.get("backend", "default")
        if backend_name == "dense":
            raise ValueError(_DENSE_BACKEND_DISABLED_MESSAGE)
        for pair in channel_pairs:
            if len(pair) != 2 or pair[0] <= 0 or pair[1] <= 0:
                raise ValueError("channel_pair must be a tuple of two positive integers")
        if backend_name == "pred_gather_igemm":
            kernel, step = _validate_pred_gather_igemm_admission(kernel_size, stride, channel_pairs, transposed=transposed)
            _validate_pred_gather_igemm_grid_admission(source_grid, target_grid)
            topology = _fvdb_cpp.gs_build_topology(source_grid.data, target_grid.data, kernel_size, stride)
            return _PredGatherIGemmBackend(gs_topology=topology, kernel_size=kernel, stride=step)
        if backend_name not in ("gather_scatter", "default"):
            raise ValueError(f"Unknown backend: {backend_name!r}")
        # identity geometry is a matmul only when row order is shared, i.e. the very same GridBatchData
        if bool((stride == 1).all()) and bool((kernel_size == 1).all()) and source_grid.data.is_same(target_grid.data):
            return _MatmulBackend()
        build = _fvdb_cpp.gs_build_transpose_topology if transposed else _fvdb_cpp.gs_build_topology
        return _GatherScatterBackend(topology=build(source_grid.data, target_grid.data, kernel_size, stride))
