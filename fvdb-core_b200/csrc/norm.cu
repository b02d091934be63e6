// norm.cu -- batch normalisation (+ fused ReLU) and column sums over jagged feature rows [N][C].
//
// The reference's fvdb.nn.BatchNorm is torch.nn.BatchNorm1d applied to jdata (fvdb/nn/modules.py:484-521), followed by
// a separate fvdb.nn.ReLU pass; for [N ~ 10^6][C <= 256] inputs torch dispatches to its NHWC kernels, which run far
// below HBM speed on this shape (profiles/r01_c3_launches.txt).  These kernels are pure streaming passes:
//   stats      1 read   per-channel (count, mean, M2) per CTA from fp32 sums, merged in double (Chan) by one small CTA
//   apply      1 read + 1 write   y = act((x - mean) * invstd * gamma + beta)
//   bwd reduce 2 reads  sum(dz), sum(dz * xhat) with dz = dy * [y > 0] recomputed from x (nothing saved but x)
//   bwd apply  2 reads + 1 write  dx = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat))
// A thread owns one 16-byte channel vector and walks rows, so every warp reads whole 128-byte lines; the
// per-channel scale / shift stay in registers.
#include "fvc_common.cuh"

namespace fvc {

constexpr int BN_THREADS = 256;
constexpr int BN_GRID = 148 * 4;
constexpr int BN_UNROLL = 1; // rows in flight per thread in the forward loops (4 measured slower on bf16: registers, not loads, limit)

template <typename T> struct RowVec;
template <> struct RowVec<float> {
    static constexpr int V = 4;
    static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct RowVec<__nv_bfloat16> {
    static constexpr int V = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { // bf16 -> fp32 is a 16-bit shift
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
template <> struct RowVec<__half> {
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

// Work split shared by every kernel: cv = C / V vectors per row; a CTA uses tpb = (BN_THREADS / cv) * cv threads, so a
// thread keeps column vector tid % cv and walks rows blockIdx * rpb + tid / cv + k * gridDim * rpb.
struct Split {
    int cv, rpb, tpb;
};
template <typename T> __host__ __device__ inline Split make_split(int c) {
    Split s;
    s.cv = c / RowVec<T>::V;
    s.rpb = BN_THREADS / s.cv;
    s.tpb = s.rpb * s.cv;
    return s;
}

// Block-level column sums of NV per-thread values: store(o, total) is called for every o = col * NV + i, col < cv.
template <int NV, typename Store>
__device__ __forceinline__ void block_column_sum(const float (&vals)[NV], const Split &sp, float *smem /*[BN_THREADS][NV + 1]*/, Store store) {
    const int tid = threadIdx.x;
    if (tid < sp.tpb) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
            smem[tid * (NV + 1) + i] = vals[i];
    }
    __syncthreads();
    for (int o = tid; o < sp.cv * NV; o += BN_THREADS) {
        const int col = o / NV, i = o % NV;
        float total = 0.f;
        for (int r = 0; r < sp.rpb; ++r)
            total += smem[(r * sp.cv + col) * (NV + 1) + i];
        store(col, i, total);
    }
}

// partial[block][2][C]: per-CTA sum and sum of squares of (x - pivot) over its rows, pivot = row 0 of the batch (shifted
// sums: with |mean| >> std the plain sum-of-squares form loses the variance to cancellation; around a pivot that is one
// sample of the distribution the cancellation is relative to ~std).  Counts follow from the row split; CTA 0 also
// publishes the pivot for the merge kernel.
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_partial_kernel(const T *__restrict__ x, int64_t n, int c, float *__restrict__ partial,
                                                                     float *__restrict__ pivot_out) {
    constexpr int V = RowVec<T>::V;
    __shared__ float smem[BN_THREADS * (2 * V + 1)];
    const Split sp = make_split<T>(c);
    const int tid = threadIdx.x;
    float acc[2 * V];
#pragma unroll
    for (int i = 0; i < 2 * V; ++i)
        acc[i] = 0.f;
    if (tid < sp.tpb) {
        const int col = tid % sp.cv;
        const int64_t step = int64_t(gridDim.x) * sp.rpb;
        float pivot[V];
        RowVec<T>::load(x + col * V, pivot); // n > 0: the host never launches an empty batch
        if (blockIdx.x == 0 && tid < sp.cv) {
#pragma unroll
            for (int i = 0; i < V; ++i)
                pivot_out[col * V + i] = pivot[i];
        }
        for (int64_t row = int64_t(blockIdx.x) * sp.rpb + tid / sp.cv; row < n; row += BN_UNROLL * step) {
            float v[BN_UNROLL][V]; // BN_UNROLL independent 16-byte loads in flight per thread
#pragma unroll
            for (int u = 0; u < BN_UNROLL; ++u) {
                if (row + u * step < n) {
                    RowVec<T>::load(x + (row + u * step) * c + col * V, v[u]);
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        v[u][i] -= pivot[i];
                } else {
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        v[u][i] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < BN_UNROLL; ++u)
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    acc[i] += v[u][i];
                    acc[V + i] = fmaf(v[u][i], v[u][i], acc[V + i]);
                }
        }
    }
    block_column_sum<2 * V>(acc, sp, smem, [&](int col, int i, float total) {
        partial[(int64_t(blockIdx.x) * 2 + (i >= V ? 1 : 0)) * c + col * V + (i % V)] = total;
    });
}

__host__ __device__ inline int64_t rows_of_block(int64_t n, int block, int grid, int rpb) {
    // rows r with (r / rpb) % grid == block
    const int64_t full = n / (int64_t(grid) * rpb), rem = n - full * grid * rpb;
    int64_t extra = rem - int64_t(block) * rpb;
    extra = extra < 0 ? 0 : (extra > rpb ? rpb : extra);
    return full * rpb + extra;
}

// Merge of the CTA partials (count, mean, M2) in double (Chan's formula): a CTA serves FIN_CH channels, 256 / FIN_CH threads
// per channel walk the partials in a fixed interleaved order, then a shared-memory tree merges them; optional
// running-stat update.
constexpr int FIN_CH = 4, FIN_LANES = 256 / FIN_CH;

__device__ __forceinline__ void chan_merge(double &cnt, double &mu, double &m2, double nb, double mb, double m2b) {
    if (nb == 0.0)
        return;
    const double delta = mb - mu, tot = cnt + nb;
    mu += delta * nb / tot;
    m2 += m2b + delta * delta * cnt * nb / tot;
    cnt = tot;
}

__global__ void __launch_bounds__(256)
bn_stats_final_kernel(const float *__restrict__ partial, const float *__restrict__ pivot, int grid, int rpb, int64_t n, int c, float *__restrict__ mean,
                      float *__restrict__ var, float *__restrict__ running_mean, float *__restrict__ running_var, float momentum) {
    __shared__ double s_cnt[256], s_mu[256], s_m2[256];
    const int tid = threadIdx.x, lane = tid / FIN_CH, ch = blockIdx.x * FIN_CH + tid % FIN_CH;
    double cnt = 0.0, mu = 0.0, m2 = 0.0;
    if (ch < c) {
        const double piv = pivot ? double(pivot[ch]) : 0.0;
        // four partials per trip: the loads are independent of the (serial) merge chain, so they overlap its latency -- the
        // convolution epilogue hands this kernel thousands of row blocks
        for (int b0 = lane; b0 < grid; b0 += 4 * FIN_LANES) {
            float s[4], ss[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int b = b0 + u * FIN_LANES;
                s[u] = b < grid ? __ldg(partial + (int64_t(b) * 2) * c + ch) : 0.f;
                ss[u] = b < grid ? __ldg(partial + (int64_t(b) * 2 + 1) * c + ch) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int b = b0 + u * FIN_LANES;
                const double nb = b < grid ? double(rows_of_block(n, b, grid, rpb)) : 0.0;
                if (nb == 0.0)
                    continue;
                const double mb = double(s[u]) / nb; // mean of (x - pivot) over the block's rows
                double m2b = double(ss[u]) - double(s[u]) * mb;
                m2b = m2b < 0.0 ? 0.0 : m2b;
                chan_merge(cnt, mu, m2, nb, mb + piv, m2b);
            }
        }
    }
    s_cnt[tid] = cnt, s_mu[tid] = mu, s_m2[tid] = m2;
    __syncthreads();
    for (int half = FIN_LANES / 2; half >= 1; half >>= 1) {
        if (lane < half) {
            const int o = tid + half * FIN_CH;
            chan_merge(s_cnt[tid], s_mu[tid], s_m2[tid], s_cnt[o], s_mu[o], s_m2[o]);
        }
        __syncthreads();
    }
    if (lane == 0 && ch < c) {
        cnt = s_cnt[tid], mu = s_mu[tid], m2 = s_m2[tid];
        const double biased = cnt > 0.0 ? m2 / cnt : 0.0;
        mean[ch] = float(mu);
        var[ch] = float(biased);
        if (running_mean)
            running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * float(mu);
        if (running_var)
            running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * float(cnt > 1.0 ? m2 / (cnt - 1.0) : biased);
    }
}

template <typename T>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const T *__restrict__ x, int64_t n, int c, const float *__restrict__ mean,
                                                              const float *__restrict__ var, const float *__restrict__ gamma,
                                                              const float *__restrict__ beta, float eps, int relu, T *__restrict__ y) {
    constexpr int V = RowVec<T>::V;
    const Split sp = make_split<T>(c);
    const int tid = threadIdx.x;
    if (tid >= sp.tpb)
        return;
    const int col = tid % sp.cv;
    float scale[V], shift[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int ch = col * V + i;
        scale[i] = rsqrtf(var[ch] + eps) * (gamma ? gamma[ch] : 1.f);
        shift[i] = (beta ? beta[ch] : 0.f) - mean[ch] * scale[i];
    }
    const int64_t step = int64_t(gridDim.x) * sp.rpb;
    for (int64_t row = int64_t(blockIdx.x) * sp.rpb + tid / sp.cv; row < n; row += BN_UNROLL * step) {
        float v[BN_UNROLL][V];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u)
            if (row + u * step < n)
                RowVec<T>::load(x + (row + u * step) * c + col * V, v[u]);
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            if (row + u * step >= n)
                break;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                v[u][i] = fmaf(v[u][i], scale[i], shift[i]);
                if (relu)
                    v[u][i] = fmaxf(v[u][i], 0.f);
            }
            RowVec<T>::store(y + (row + u * step) * c + col * V, v[u]);
        }
    }
}

// partial[block][2][C]: sum(dz), sum(dz * xhat); dz = dy * [act input > 0] when relu
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_backward_reduce_kernel(const T *__restrict__ dy, const T *__restrict__ x, int64_t n, int c, const float *__restrict__ mean,
                          const float *__restrict__ var, const float *__restrict__ gamma, const float *__restrict__ beta, float eps, int relu,
                          float *__restrict__ partial) {
    constexpr int V = RowVec<T>::V;
    __shared__ float smem[BN_THREADS * (2 * V + 1)];
    const Split sp = make_split<T>(c);
    const int tid = threadIdx.x;
    float acc[2 * V];
#pragma unroll
    for (int i = 0; i < 2 * V; ++i)
        acc[i] = 0.f;
    if (tid < sp.tpb) {
        const int col = tid % sp.cv;
        float mu[V], inv[V], g[V], b[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int ch = col * V + i;
            mu[i] = mean[ch], inv[i] = rsqrtf(var[ch] + eps), g[i] = gamma ? gamma[ch] : 1.f, b[i] = beta ? beta[ch] : 0.f;
        }
        const int64_t step = int64_t(gridDim.x) * sp.rpb;
        for (int64_t row = int64_t(blockIdx.x) * sp.rpb + tid / sp.cv; row < n; row += 2 * step) {
            float xv[2][V], dv[2][V]; // two rows (four loads) in flight per thread
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (row + u * step < n) {
                    RowVec<T>::load(x + (row + u * step) * c + col * V, xv[u]);
                    RowVec<T>::load(dy + (row + u * step) * c + col * V, dv[u]);
                } else {
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        xv[u][i] = mu[i], dv[u][i] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float xhat = (xv[u][i] - mu[i]) * inv[i];
                    const float dz = (relu && fmaf(xhat, g[i], b[i]) <= 0.f) ? 0.f : dv[u][i];
                    acc[i] += dz;
                    acc[V + i] = fmaf(dz, xhat, acc[V + i]);
                }
        }
    }
    block_column_sum<2 * V>(acc, sp, smem, [&](int col, int i, float total) {
        partial[(int64_t(blockIdx.x) * 2 + (i >= V ? 1 : 0)) * c + col * V + (i % V)] = total;
    });
}

// sums[plane][C] = sum over blocks of partial[.][plane][C] (double accumulation, fixed order): FIN_LANES threads per
// output walk the partials interleaved, then a shared-memory tree
__global__ void __launch_bounds__(256) column_sums_final_kernel(const float *__restrict__ partial, int grid, int c, int planes, float *__restrict__ sums) {
    __shared__ double s_tot[256];
    const int tid = threadIdx.x, lane = tid / FIN_CH, e = blockIdx.x * FIN_CH + tid % FIN_CH;
    double total = 0.0;
    if (e < planes * c) {
        const int plane = e / c, ch = e % c;
        for (int b = lane; b < grid; b += FIN_LANES)
            total += partial[(int64_t(b) * planes + plane) * c + ch];
    }
    s_tot[tid] = total;
    __syncthreads();
    for (int half = FIN_LANES / 2; half >= 1; half >>= 1) {
        if (lane < half)
            s_tot[tid] += s_tot[tid + half * FIN_CH];
        __syncthreads();
    }
    if (lane == 0 && e < planes * c)
        sums[e] = float(s_tot[tid]);
}

// dx = gamma * invstd * (dz - sum_dz / count - xhat * sum_dz_xhat / count)   (training)
// dx = gamma * invstd * dz                                                  (eval: statistics are constants)
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_backward_apply_kernel(const T *__restrict__ dy, const T *__restrict__ x, int64_t n, int c, const float *__restrict__ mean,
                         const float *__restrict__ var, const float *__restrict__ gamma, const float *__restrict__ beta, float eps, int relu,
                         int training, const float *__restrict__ sums /*[2][C]*/, float inv_count, const float *__restrict__ count_dev,
                         T *__restrict__ dx) {
    constexpr int V = RowVec<T>::V;
    const Split sp = make_split<T>(c);
    const int tid = threadIdx.x;
    if (tid >= sp.tpb)
        return;
    const int col = tid % sp.cv;
    float mu[V], inv[V], g[V], b[V], m_dz[V], m_dzx[V];
    if (count_dev) // total row count of a distributed batch, kept on the device (no host round trip)
        inv_count = 1.f / fmaxf(__ldg(count_dev), 1.f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int ch = col * V + i;
        mu[i] = mean[ch], inv[i] = rsqrtf(var[ch] + eps), g[i] = gamma ? gamma[ch] : 1.f, b[i] = beta ? beta[ch] : 0.f;
        m_dz[i] = training ? sums[ch] * inv_count : 0.f;
        m_dzx[i] = training ? sums[c + ch] * inv_count : 0.f;
    }
    const int64_t step = int64_t(gridDim.x) * sp.rpb;
    for (int64_t row = int64_t(blockIdx.x) * sp.rpb + tid / sp.cv; row < n; row += 2 * step) {
        float xv[2][V], dv[2][V];
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (row + u * step < n) {
                RowVec<T>::load(x + (row + u * step) * c + col * V, xv[u]);
                RowVec<T>::load(dy + (row + u * step) * c + col * V, dv[u]);
            }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (row + u * step >= n)
                break;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float xhat = (xv[u][i] - mu[i]) * inv[i];
                const float dz = (relu && fmaf(xhat, g[i], b[i]) <= 0.f) ? 0.f : dv[u][i];
                dv[u][i] = g[i] * inv[i] * (dz - m_dz[i] - xhat * m_dzx[i]);
            }
            RowVec<T>::store(dx + (row + u * step) * c + col * V, dv[u]);
        }
    }
}

// partial[block][1][C] = column sums of the CTA's rows (bias gradient)
template <typename T>
__global__ void __launch_bounds__(BN_THREADS) column_sums_partial_kernel(const T *__restrict__ x, int64_t n, int c, float *__restrict__ partial) {
    constexpr int V = RowVec<T>::V;
    __shared__ float smem[BN_THREADS * (V + 1)];
    const Split sp = make_split<T>(c);
    const int tid = threadIdx.x;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i)
        acc[i] = 0.f;
    if (tid < sp.tpb) {
        const int col = tid % sp.cv;
        for (int64_t row = int64_t(blockIdx.x) * sp.rpb + tid / sp.cv; row < n; row += int64_t(gridDim.x) * sp.rpb) {
            float v[V];
            RowVec<T>::load(x + row * c + col * V, v);
#pragma unroll
            for (int i = 0; i < V; ++i)
                acc[i] += v[i];
        }
    }
    block_column_sum<V>(acc, sp, smem, [&](int col, int i, float total) { partial[int64_t(blockIdx.x) * c + col * V + i] = total; });
}

static int bn_grid(int64_t n, int rpb) {
    const int64_t blocks = ceil_div(n > 0 ? n : 1, rpb);
    return int(blocks < BN_GRID ? blocks : BN_GRID);
}

static int check_rows(const char *name, int64_t n, int32_t c, int32_t dtype) {
    FVC_REQUIRE(dtype == FVC_F16 || dtype == FVC_BF16 || dtype == FVC_F32, FVC_ERR_UNSUPPORTED, "%s: dtype code %d is not served (f16, bf16, f32)", name, dtype);
    const int v = dtype == FVC_F32 ? 4 : 8;
    FVC_REQUIRE(c > 0 && c % v == 0 && c / v <= BN_THREADS, FVC_ERR_UNSUPPORTED, "%s: channel count %d must be a multiple of %d and <= %d", name, c, v, v * BN_THREADS);
    FVC_REQUIRE(n >= 0, FVC_ERR_VALUE, "%s: negative row count", name);
    return FVC_OK;
}

#define FVC_BY_DTYPE(dtype, ...)                                   \
    switch (dtype) {                                               \
    case FVC_F16: { using T = __half; __VA_ARGS__; } break;         \
    case FVC_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
    default: { using T = float; __VA_ARGS__; } break;               \
    }

} // namespace fvc

using namespace fvc;

extern "C" {

// [BN_GRID][2][C] CTA partials | [C] pivot row of the statistics pass
size_t fvc_bn_scratch_bytes(int32_t channels) { return (size_t(BN_GRID) * 2 + 1) * size_t(channels > 0 ? channels : 0) * sizeof(float) + 256; }

int fvc_bn_stats(const void *x, int64_t n, int32_t c, int32_t dtype, float *mean, float *var, float *running_mean, float *running_var,
                 float momentum, void *scratch, size_t scratch_bytes, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_rows("fvc_bn_stats", n, c, dtype);
    if (rc)
        return rc;
    FVC_REQUIRE(mean && var && scratch && scratch_bytes >= fvc_bn_scratch_bytes(c), FVC_ERR_RUNTIME, "fvc_bn_stats: null output or scratch too small");
    FVC_REQUIRE(n == 0 || x, FVC_ERR_RUNTIME, "fvc_bn_stats: null input");
    float *partial = reinterpret_cast<float *>(scratch);
    float *pivot = partial + size_t(BN_GRID) * 2 * size_t(c);
    int grid = 1, rpb = 1;
    if (n == 0) // nothing to launch: the merge kernel writes mean = var = 0 from zero partial counts
        FVC_CUDA(cudaMemsetAsync(pivot, 0, size_t(c) * sizeof(float), stream));
    FVC_BY_DTYPE(dtype, {
        const Split sp = make_split<T>(c);
        rpb = sp.rpb;
        grid = bn_grid(n, rpb);
        if (n > 0)
            bn_stats_partial_kernel<T><<<grid, BN_THREADS, 0, stream>>>(reinterpret_cast<const T *>(x), n, c, partial, pivot);
    });
    FVC_LAUNCH_CHECK();
    bn_stats_final_kernel<<<int(ceil_div(c, FIN_CH)), 256, 0, stream>>>(partial, pivot, grid, rpb, n, c, mean, var, running_mean, running_var, momentum);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_bn_stats_from_partials(const float *partial, int64_t blocks, int32_t rows_per_block, int64_t n, int32_t c, float *mean, float *var,
                               float *running_mean, float *running_var, float momentum, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(c > 0 && mean && var, FVC_ERR_RUNTIME, "fvc_bn_stats_from_partials: null output");
    FVC_REQUIRE(blocks >= 0 && blocks <= INT32_MAX && rows_per_block > 0 && (blocks == 0 || partial), FVC_ERR_RUNTIME,
                "fvc_bn_stats_from_partials: bad partial layout");
    FVC_REQUIRE(n <= blocks * int64_t(rows_per_block) && n > (blocks - 1) * int64_t(rows_per_block), FVC_ERR_RUNTIME,
                "fvc_bn_stats_from_partials: %lld blocks of %d rows do not cover %lld rows", (long long)blocks, rows_per_block, (long long)n);
    // block b holds rows [b * rpb, (b + 1) * rpb): with grid == blocks that is exactly the interleaved split of the merge kernel
    bn_stats_final_kernel<<<int(ceil_div(c, FIN_CH)), 256, 0, stream>>>(partial, nullptr, int(blocks), rows_per_block, n, c, mean, var, running_mean,
                                                                       running_var, momentum);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_bn_apply(const void *x, int64_t n, int32_t c, int32_t dtype, const float *mean, const float *var, const float *gamma, const float *beta,
                 float eps, int32_t relu, void *y, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_rows("fvc_bn_apply", n, c, dtype);
    if (rc)
        return rc;
    if (n == 0)
        return FVC_OK;
    FVC_REQUIRE(x && y && mean && var, FVC_ERR_RUNTIME, "fvc_bn_apply: null pointer");
    FVC_BY_DTYPE(dtype, {
        const Split sp = make_split<T>(c);
        bn_apply_kernel<T><<<bn_grid(n, sp.rpb), BN_THREADS, 0, stream>>>(reinterpret_cast<const T *>(x), n, c, mean, var, gamma, beta, eps, relu,
                                                                         reinterpret_cast<T *>(y));
    });
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_bn_backward_reduce(const void *dy, const void *x, int64_t n, int32_t c, int32_t dtype, const float *mean, const float *var,
                           const float *gamma, const float *beta, float eps, int32_t relu, float *sums, void *scratch, size_t scratch_bytes,
                           fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_rows("fvc_bn_backward_reduce", n, c, dtype);
    if (rc)
        return rc;
    FVC_REQUIRE(sums && mean && var && scratch && scratch_bytes >= fvc_bn_scratch_bytes(c), FVC_ERR_RUNTIME, "fvc_bn_backward_reduce: null output or scratch too small");
    FVC_REQUIRE(n == 0 || (x && dy), FVC_ERR_RUNTIME, "fvc_bn_backward_reduce: null input");
    float *partial = reinterpret_cast<float *>(scratch);
    int grid = 1;
    FVC_BY_DTYPE(dtype, {
        const Split sp = make_split<T>(c);
        grid = bn_grid(n, sp.rpb);
        bn_backward_reduce_kernel<T><<<grid, BN_THREADS, 0, stream>>>(reinterpret_cast<const T *>(dy), reinterpret_cast<const T *>(x), n, c, mean, var,
                                                                     gamma, beta, eps, relu, partial);
    });
    FVC_LAUNCH_CHECK();
    column_sums_final_kernel<<<int(ceil_div(2 * c, FIN_CH)), 256, 0, stream>>>(partial, grid, c, 2, sums);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_bn_backward_apply(const void *dy, const void *x, int64_t n, int32_t c, int32_t dtype, const float *mean, const float *var,
                          const float *gamma, const float *beta, float eps, int32_t relu, int32_t training, const float *sums, int64_t count,
                          const float *count_dev, void *dx, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_rows("fvc_bn_backward_apply", n, c, dtype);
    if (rc)
        return rc;
    if (n == 0)
        return FVC_OK;
    FVC_REQUIRE(x && dy && dx && mean && var && (!training || sums), FVC_ERR_RUNTIME, "fvc_bn_backward_apply: null pointer");
    const float inv_count = count > 0 ? 1.f / float(count) : 0.f;
    FVC_BY_DTYPE(dtype, {
        const Split sp = make_split<T>(c);
        bn_backward_apply_kernel<T><<<bn_grid(n, sp.rpb), BN_THREADS, 0, stream>>>(reinterpret_cast<const T *>(dy), reinterpret_cast<const T *>(x), n, c,
                                                                                  mean, var, gamma, beta, eps, relu, training, sums, inv_count,
                                                                                  count_dev, reinterpret_cast<T *>(dx));
    });
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int fvc_column_sums(const void *x, int64_t n, int32_t c, int32_t dtype, float *sums, void *scratch, size_t scratch_bytes, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int rc = check_rows("fvc_column_sums", n, c, dtype);
    if (rc)
        return rc;
    FVC_REQUIRE(sums && scratch && scratch_bytes >= fvc_bn_scratch_bytes(c), FVC_ERR_RUNTIME, "fvc_column_sums: null output or scratch too small");
    FVC_REQUIRE(n == 0 || x, FVC_ERR_RUNTIME, "fvc_column_sums: null input");
    float *partial = reinterpret_cast<float *>(scratch);
    int grid = 1;
    FVC_BY_DTYPE(dtype, {
        const Split sp = make_split<T>(c);
        grid = bn_grid(n, sp.rpb);
        column_sums_partial_kernel<T><<<grid, BN_THREADS, 0, stream>>>(reinterpret_cast<const T *>(x), n, c, partial);
    });
    FVC_LAUNCH_CHECK();
    column_sums_final_kernel<<<int(ceil_div(c, FIN_CH)), 256, 0, stream>>>(partial, grid, c, 1, sums);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

} // extern "C"
