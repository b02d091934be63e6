// fvc_common.cuh -- shared host/device helpers for libfvdbconv (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/fvdbconv.h"

static_assert(sizeof(FvcLeaf) == 128, "FvcLeaf must be exactly one 128-byte line");

namespace fvc {

// ---- error plumbing ---------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);
extern std::atomic<int64_t> g_launch_count;

#define FVC_REQUIRE(cond, code, ...)                  \
    do {                                              \
        if (!(cond))                                  \
            return ::fvc::set_error(code, __VA_ARGS__); \
    } while (0)

#define FVC_CUDA(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t err__ = (expr);                                                                          \
        if (err__ != cudaSuccess)                                                                            \
            return ::fvc::set_error(FVC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                                    __FILE__, __LINE__);                                                     \
    } while (0)

// every kernel launch goes through this so that bench.py can report gpu_launches
#define FVC_LAUNCH_CHECK()                  \
    do {                                    \
        ::fvc::g_launch_count.fetch_add(1); \
        FVC_CUDA(cudaGetLastError());       \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- geometry (ConvolutionGeometry.h:85-146,174-177), passed to kernels by value --------------
struct Geometry {
    int32_t k[3];
    int32_t s[3];
    int32_t pad[3];
    int64_t volume;
};

inline Geometry make_geometry(const int32_t kernel_size[3], const int32_t stride[3]) {
    Geometry g;
    g.volume = 1;
    for (int d = 0; d < 3; ++d) {
        g.k[d] = kernel_size[d];
        g.s[d] = stride[d];
        g.pad[d] = (kernel_size[d] - 1) / 2;
        g.volume *= kernel_size[d];
    }
    return g;
}

__host__ __device__ __forceinline__ int32_t floor_div(int32_t a, int32_t b) {
    int32_t q = a / b, r = a % b;
    return r < 0 ? q - 1 : q;
}
__host__ __device__ __forceinline__ int32_t floor_mod(int32_t a, int32_t b) {
    int32_t r = a % b;
    return r < 0 ? r + b : r;
}

// ---- index-grid lookups -----------------------------------------------------------------------
// Root tile search: a grid owns a handful of 4096^3 tiles; linear scan over its (sorted) range.
__device__ __forceinline__ int find_upper(const FvcGridBatch &g, int b, int x, int y, int z) {
    const int tx = x >> 12, ty = y >> 12, tz = z >> 12;
    const int lo = __ldg(g.root_offsets + b), hi = __ldg(g.root_offsets + b + 1);
    const int4 *keys = reinterpret_cast<const int4 *>(g.root_keys);
    for (int r = lo; r < hi; ++r) {
        int4 key = __ldg(keys + r);
        if (key.y == tx && key.z == ty && key.w == tz)
            return r;
    }
    return -1;
}

// Leaf index containing (x,y,z) in grid b, or -1.  Four dependent loads: root, upper, lower.
__device__ __forceinline__ int find_leaf(const FvcGridBatch &g, int b, int x, int y, int z) {
    const int up = find_upper(g, b, x, y, z);
    if (up < 0)
        return -1;
    const int uo = ((((x >> 7) & 31) << 5) | ((y >> 7) & 31)) << 5 | ((z >> 7) & 31);
    const int low = __ldg(g.upper + (int64_t(up) << 15) + uo);
    if (low < 0)
        return -1;
    const int lo = ((((x >> 3) & 15) << 4) | ((y >> 3) & 15)) << 4 | ((z >> 3) & 15);
    return __ldg(g.lower + (int64_t(low) << 12) + lo);
}

// 0-based batch-cumulative row of (x,y,z) inside a leaf given its mask word / prefix / base, or -1.
__device__ __forceinline__ int leaf_value(uint64_t mask_word, uint32_t prefix, int base, int y, int z) {
    const int bit = ((y & 7) << 3) | (z & 7);
    if (!((mask_word >> bit) & 1ull))
        return -1;
    return base + int(prefix) + __popcll(mask_word & ((1ull << bit) - 1ull));
}

__device__ __forceinline__ int lookup_row(const FvcGridBatch &g, int b, int x, int y, int z) {
    const int leaf = find_leaf(g, b, x, y, z);
    if (leaf < 0)
        return -1;
    const FvcLeaf *L = g.leaves + leaf;
    const int w = x & 7;
    return leaf_value(__ldg(L->mask + w), __ldg(L->prefix + w), __ldg(&L->base), y, z);
}

// ---- dtype helpers ----------------------------------------------------------------------------
template <typename T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };

template <typename A, typename T> __device__ __forceinline__ A to_acc(T v) { return static_cast<A>(v); }
template <> __device__ __forceinline__ float to_acc<float, __half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_acc<float, __nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T, typename A> __device__ __forceinline__ T from_acc(A v) { return static_cast<T>(v); }
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16, float>(float v) { return __float2bfloat16_rn(v); }

inline size_t dtype_size(int dtype) {
    switch (dtype) {
    case FVC_F16:
    case FVC_BF16: return 2;
    case FVC_F32: return 4;
    case FVC_F64: return 8;
    default: return 0;
    }
}

} // namespace fvc
