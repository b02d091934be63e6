// pool.cu -- row pooling / row gathers between a fine and a coarse grid: the data movers behind GridBatch.max_pool,
// avg_pool and refine (reference: ops/MaxPool.cu:16-122, ops/AvgPool.cu:17-110, ops/Refine.cu:17-110), SURVEY.md 8(f) rank 3.
//
// The reference walks the NanoVDB tree once per (coarse voxel, window tap, CHANNEL).  Here the window lookups are done
// once per (coarse voxel, tap) by the index-grid lookup kernel (fvc_ijk_to_index) into a child table idx[n_out][taps];
// these kernels then stream rows: a thread owns one 16-byte channel vector of one output row and reduces its <= taps
// children, so every access is a whole-row vector load.  No atomics: every destination element has exactly one writer
// (non-overlapping windows), as in the reference's own backward passes.
#include "fvc_common.cuh"

namespace fvc {

constexpr int POOL_THREADS = 256;
enum { POOL_MAX = 0, POOL_SUM = 1 };

template <typename T> struct PoolVec;
template <> struct PoolVec<float> {
    static constexpr int V = 4;
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct PoolVec<__nv_bfloat16> {
    static constexpr int V = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct PoolVec<__half> {
    static constexpr int V = 8;
    static __device__ __forceinline__ void load(const __half *p, float (&v)[8]) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
            v[2 * i] = f.x, v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__half *p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// y[o] = reduce over taps t with idx[o][t] >= 0 of x[idx[o][t]]  (sum: scaled).  A window without any active child yields
// ZERO: that is the documented contract of MaxPool (fvdb/nn/modules.py:125-128 "not covered by any source voxels ... set to
// zero"); the reference kernel itself leaves its -INFINITY initialiser there (MaxPool.cu:46), which poisons every network
// that pools onto a dilated coarse grid (fvdb/nn/simple_unet.py:433).
template <typename T>
__global__ void __launch_bounds__(POOL_THREADS)
pool_rows_kernel(const T *__restrict__ x, const int32_t *__restrict__ idx, int64_t n_out, int taps, int c, int mode, float scale,
                 T *__restrict__ y) {
    constexpr int V = PoolVec<T>::V;
    const int cv = c / V;
    const int64_t total = n_out * cv;
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t o = e / cv;
        const int col = int(e - o * cv) * V;
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i)
            acc[i] = mode == POOL_MAX ? -INFINITY : 0.f;
        bool any = false;
        for (int t = 0; t < taps; ++t) {
            const int r = __ldg(idx + o * taps + t);
            if (r < 0)
                continue;
            any = true;
            float v[V];
            PoolVec<T>::load(x + int64_t(r) * c + col, v);
#pragma unroll
            for (int i = 0; i < V; ++i)
                acc[i] = mode == POOL_MAX ? fmaxf(acc[i], v[i]) : acc[i] + v[i];
        }
        if (!any) {
#pragma unroll
            for (int i = 0; i < V; ++i)
                acc[i] = 0.f;
        }
        if (mode == POOL_SUM) {
#pragma unroll
            for (int i = 0; i < V; ++i)
                acc[i] *= scale;
        }
        PoolVec<T>::store(y + o * c + col, acc);
    }
}

// max: dx[idx[o][argmax_t x[idx[o][t]][ch]]][ch] = dy[o][ch] (first maximum wins, MaxPool.cu:100-118);
// sum: dx[idx[o][t]][ch] = dy[o][ch] * scale for every child (AvgPool.cu:98-108).  dx is zero-initialised by the caller.
template <typename T>
__global__ void __launch_bounds__(POOL_THREADS)
pool_rows_backward_kernel(const T *__restrict__ dy, const T *__restrict__ x, const int32_t *__restrict__ idx, int64_t n_out, int taps, int c,
                          int mode, float scale, T *__restrict__ dx) {
    constexpr int V = PoolVec<T>::V;
    const int cv = c / V;
    const int64_t total = n_out * cv;
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t o = e / cv;
        const int col = int(e - o * cv) * V;
        float g[V];
        PoolVec<T>::load(dy + o * c + col, g);
        if (mode == POOL_SUM) {
#pragma unroll
            for (int i = 0; i < V; ++i)
                g[i] *= scale;
            for (int t = 0; t < taps; ++t) {
                const int r = __ldg(idx + o * taps + t);
                if (r >= 0)
                    PoolVec<T>::store(dx + int64_t(r) * c + col, g);
            }
            continue;
        }
        float best[V];
        int arg[V];
#pragma unroll
        for (int i = 0; i < V; ++i)
            best[i] = -INFINITY, arg[i] = -1;
        for (int t = 0; t < taps; ++t) {
            const int r = __ldg(idx + o * taps + t);
            if (r < 0)
                continue;
            float v[V];
            PoolVec<T>::load(x + int64_t(r) * c + col, v);
#pragma unroll
            for (int i = 0; i < V; ++i)
                if (v[i] > best[i])
                    best[i] = v[i], arg[i] = r;
        }
        // children of one window are distinct rows, and windows do not overlap: element-wise scalar stores, one writer each
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (arg[i] >= 0)
                dx[int64_t(arg[i]) * c + col + i] = from_acc<T, float>(g[i]);
    }
}

// y[r] = idx[r] >= 0 ? x[idx[r]] : 0   (nearest-neighbour refine, Refine.cu:43-56)
template <typename T>
__global__ void __launch_bounds__(POOL_THREADS)
gather_rows_kernel(const T *__restrict__ x, const int32_t *__restrict__ idx, int64_t n_out, int c, T *__restrict__ y) {
    constexpr int V = PoolVec<T>::V;
    const int cv = c / V;
    const int64_t total = n_out * cv;
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t o = e / cv;
        const int col = int(e - o * cv) * V;
        const int r = __ldg(idx + o);
        float v[V];
#pragma unroll
        for (int i = 0; i < V; ++i)
            v[i] = 0.f;
        if (r >= 0)
            PoolVec<T>::load(x + int64_t(r) * c + col, v);
        PoolVec<T>::store(y + o * c + col, v);
    }
}

static int check_pool(const char *name, int64_t n_out, int32_t taps, int32_t c, int32_t dtype) {
    FVC_REQUIRE(dtype == FVC_F16 || dtype == FVC_BF16 || dtype == FVC_F32, FVC_ERR_UNSUPPORTED, "%s: dtype code %d is not served (f16, bf16, f32)", name, dtype);
    const int v = dtype == FVC_F32 ? 4 : 8;
    FVC_REQUIRE(c > 0 && c % v == 0, FVC_ERR_UNSUPPORTED, "%s: channel count %d must be a multiple of %d", name, c, v);
    FVC_REQUIRE(n_out >= 0 && taps >= 0, FVC_ERR_VALUE, "%s: negative size", name);
    return FVC_OK;
}

static int pool_grid(int64_t vectors) {
    const int64_t blocks = ceil_div(vectors > 0 ? vectors : 1, POOL_THREADS);
    return int(blocks < 148 * 16 ? blocks : 148 * 16);
}

#define FVC_POOL_BY_DTYPE(dtype, ...)                               \
    switch (dtype) {                                               \
    case FVC_F16: { using T = __half; __VA_ARGS__; } break;         \
    case FVC_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
    default: { using T = float; __VA_ARGS__; } break;               \
    }

} // namespace fvc

using namespace fvc;

extern "C" {

int fvc_pool_rows(const void *x, const int32_t *idx, int64_t n_out, int32_t taps, int32_t channels, int32_t dtype, int32_t mode, float scale,
                  void *y, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_pool("fvc_pool_rows", n_out, taps, channels, dtype);
    if (rc)
        return rc;
    FVC_REQUIRE(mode == POOL_MAX || mode == POOL_SUM, FVC_ERR_VALUE, "fvc_pool_rows: mode must be 0 (max) or 1 (scaled sum)");
    if (n_out == 0)
        return FVC_OK;
    FVC_REQUIRE(y && (taps == 0 || idx), FVC_ERR_RUNTIME, "fvc_pool_rows: null pointer");
    const int v = dtype == FVC_F32 ? 4 : 8;
    FVC_POOL_BY_DTYPE(dtype, pool_rows_kernel<T><<<pool_grid(n_out * (channels / v)), POOL_THREADS, 0, stream>>>(
                                 reinterpret_cast<const T *>(x), idx, n_out, taps, channels, mode, scale, reinterpret_cast<T *>(y)));
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_pool_rows_backward(const void *dy, const void *x, const int32_t *idx, int64_t n_out, int32_t taps, int32_t channels, int32_t dtype,
                           int32_t mode, float scale, void *dx, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_pool("fvc_pool_rows_backward", n_out, taps, channels, dtype);
    if (rc)
        return rc;
    FVC_REQUIRE(mode == POOL_MAX || mode == POOL_SUM, FVC_ERR_VALUE, "fvc_pool_rows_backward: mode must be 0 (max) or 1 (scaled sum)");
    if (n_out == 0 || taps == 0)
        return FVC_OK;
    FVC_REQUIRE(dy && dx && idx && (mode == POOL_SUM || x), FVC_ERR_RUNTIME, "fvc_pool_rows_backward: null pointer");
    const int v = dtype == FVC_F32 ? 4 : 8;
    FVC_POOL_BY_DTYPE(dtype, pool_rows_backward_kernel<T><<<pool_grid(n_out * (channels / v)), POOL_THREADS, 0, stream>>>(
                                 reinterpret_cast<const T *>(dy), reinterpret_cast<const T *>(x), idx, n_out, taps, channels, mode, scale,
                                 reinterpret_cast<T *>(dx)));
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_gather_rows(const void *x, const int32_t *idx, int64_t n_out, int32_t channels, int32_t dtype, void *y, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_pool("fvc_gather_rows", n_out, 1, channels, dtype);
    if (rc)
        return rc;
    if (n_out == 0)
        return FVC_OK;
    FVC_REQUIRE(y && idx, FVC_ERR_RUNTIME, "fvc_gather_rows: null pointer");
    const int v = dtype == FVC_F32 ? 4 : 8;
    FVC_POOL_BY_DTYPE(dtype, gather_rows_kernel<T><<<pool_grid(n_out * (channels / v)), POOL_THREADS, 0, stream>>>(
                                 reinterpret_cast<const T *>(x), idx, n_out, channels, reinterpret_cast<T *>(y)));
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

} // extern "C"
