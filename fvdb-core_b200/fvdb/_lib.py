"""ctypes binding of libfvdbconv.so (include/fvdbconv.h).  No CPU fallback: import fails loudly if the
library has not been built, and every compute entry point fails without a CUDA device."""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libfvdbconv.so"

FVC_OK, FVC_ERR_VALUE, FVC_ERR_RUNTIME, FVC_ERR_INDEX, FVC_ERR_CUDA, FVC_ERR_UNSUPPORTED = range(6)
FVC_F16, FVC_BF16, FVC_F32, FVC_F64 = range(4)
ABI_VERSION = 3


class FvcGridBatch(C.Structure):
    _fields_ = [
        ("num_grids", C.c_int32),
        ("num_leaves", C.c_int32),
        ("num_lower", C.c_int32),
        ("num_upper", C.c_int32),
        ("total_voxels", C.c_int64),
        ("leaves", C.c_void_p),
        ("lower", C.c_void_p),
        ("upper", C.c_void_p),
        ("root_keys", C.c_void_p),
        ("root_offsets", C.c_void_p),
        ("voxel_offsets", C.c_void_p),
        ("leaf_offsets", C.c_void_p),
    ]


class FvcConvEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("residual", C.c_void_p), ("relu", C.c_int32), ("stats", C.c_void_p)]


_I3 = C.c_int32 * 3
_vp, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
_GB = C.POINTER(FvcGridBatch)

# name -> (restype, argtypes); must list every symbol include/fvdbconv.h declares
SIGNATURES = {
    "fvc_abi_version": (C.c_int, []),
    "fvc_last_error": (C.c_char_p, []),
    "fvc_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "fvc_launch_count": (_i64, []),
    "fvc_geometry": (C.c_int, [_I3, _I3, _I3, _I3, C.POINTER(_i64)]),
    "fvc_geometry_tap_coord": (C.c_int, [_I3, _i64, _I3]),
    "fvc_geometry_fine_from_coarse": (C.c_int, [_I3, _I3, _I3, _I3, _I3]),
    "fvc_geometry_coarse_from_fine": (C.c_int, [_I3, _I3, _I3, _I3, _I3, C.POINTER(_i32)]),
    "fvc_grid_build_scratch_bytes": (_sz, [_i64]),
    "fvc_grid_build_count": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _sz, C.POINTER(_i64 * 4), _vp]),
    "fvc_grid_build_fill": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _sz, C.POINTER(_i64 * 4)] + [_vp] * 9 + [_vp]),
    "fvc_conv_grid_count": (C.c_int, [_vp, _i64, _I3, _I3, _i32, _vp, C.POINTER(_i64), _vp]),
    "fvc_conv_grid_emit": (C.c_int, [_vp, _vp, _i64, _I3, _I3, _i32, _i64, _vp, _vp, _vp, _vp]),
    "fvc_grid_dilate_leaves": (C.c_int, [_GB, _vp, _i32, _I3, _I3, _vp, _vp]),
    "fvc_grid_expand_leaves": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "fvc_kmap_build": (C.c_int, [_GB, _GB, _I3, _I3, _i32, _vp, _i64, _vp, _vp, _vp]),
    "fvc_kmap_reverse_from_dense": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp]),
    "fvc_kmap_csr_scratch_bytes": (_sz, [_i64, _i64]),
    "fvc_kmap_to_csr": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "fvc_kmap_reverse_dense": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "fvc_kmap_tile_mask": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "fvc_kmap_degree": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "fvc_neighbor_indexes": (C.c_int, [_GB, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "fvc_ijk_to_index": (C.c_int, [_GB, _vp, _vp, _i64, _i32, _vp, _vp]),
    "fvc_pack_weights": (C.c_int, [_vp, C.POINTER(_i64 * 5), _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "fvc_conv_scratch_bytes": (_sz, [_i64, _i64, _i32, _i32, _i64, _i32]),
    "fvc_conv_forward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _vp, _sz, _vp]),
    "fvc_conv_wgrad_scratch_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _i32]),
    "fvc_set_tuning": (C.c_int, [_i32, _i32]),
    "fvc_conv_backward_fused_scratch_bytes": (_sz, [_i64, _i32, _i32, _i64]),
    "fvc_conv_backward_fused": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "fvc_conv_kernel_family": (_i32, [_i32, _i32, _i64, _i32, _i32, _i32]),
    "fvc_conv_weights_bytes": (_sz, [_i32, _i32, _i64, _i32, _i32]),
    "fvc_conv_prepare_weights": (C.c_int, [_vp, C.POINTER(_i64 * 5), _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _sz, _vp]),
    "fvc_conv_stats_blocks": (_i64, [_i64, _i32, _i32, _i64, _i32, _i32, C.POINTER(_i32)]),
    "fvc_conv_forward_ex": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _vp, _sz, _vp]),
    "fvc_split_rows": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "fvc_conv_wgrad_ex": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, C.POINTER(_i64), _vp, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fvc_bn_stats_from_partials": (C.c_int, [_vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _vp, C.c_float, _vp]),
    "fvc_bn_scratch_bytes": (_sz, [_i32]),
    "fvc_bn_stats": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, C.c_float, _vp, _sz, _vp]),
    "fvc_bn_apply": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, C.c_float, _i32, _vp, _vp]),
    "fvc_bn_backward_reduce": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, C.c_float, _i32, _vp, _vp, _sz, _vp]),
    "fvc_bn_backward_apply": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, C.c_float, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "fvc_column_sums": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "fvc_pool_rows": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, C.c_float, _vp, _vp]),
    "fvc_pool_rows_backward": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, C.c_float, _vp, _vp]),
    "fvc_gather_rows": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "fvc_conv_wgrad": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(_i64), _vp, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
}


def _load() -> C.CDLL:
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python fvdb-core_b200/build.py` (nvcc, sm_100a). "
            "There is no CPU fallback for the convolution path."
        )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.fvc_abi_version() != ABI_VERSION:
        raise ImportError(f"libfvdbconv ABI version {lib.fvc_abi_version()} != expected {ABI_VERSION}")
    return lib


lib = _load()


def last_error() -> str:
    return lib.fvc_last_error().decode(errors="replace")


def check(rc: int) -> None:
    """Raise the exception class the reference's TORCH_CHECK* family would."""
    if rc == FVC_OK:
        return
    msg = last_error()
    if rc == FVC_ERR_VALUE:
        raise ValueError(msg)
    if rc == FVC_ERR_INDEX:
        raise IndexError(msg)
    raise RuntimeError(msg)


def i3(values) -> "_I3":
    return _I3(*[int(v) for v in values])


def launch_count() -> int:
    return int(lib.fvc_launch_count())
