// conv_simt.cu -- CUDA-core (FFMA / DFMA) sparse convolution: forward / dgrad (output-stationary over the
// dense tap-major map, no atomics) and weight gradient (tap-stationary over the CSR pairs, fixed-order
// reduction).  Serves fp32 (true fp32, the 1e-5 parity bar), fp64 (the reference's 1e-11 tests), and any
// channel count the tensor-core path does not admit.  Replaces the per-tap gather -> torch::mm ->
// atomic scatter-add pipeline of GatherScatterDefault.cu:536-586,673-816.
#include "conv_internal.cuh"

namespace fvc {

constexpr int SIMT_THREADS = 256;
constexpr int SIMT_KC = 16; // reduction (channel / pair) chunk staged per step

// ---- forward / dgrad ------------------------------------------------------------------------------
// CTA tile: TM output rows x TN output channels, thread micro-tile 4 x 4; reduction over (tap, Cin).
template <typename T, int TN>
__global__ void __launch_bounds__(SIMT_THREADS)
conv_os_simt_kernel(const T *__restrict__ x, const T *__restrict__ w, const T *__restrict__ bias, T *__restrict__ y,
                    const int32_t *__restrict__ nbr, int64_t pitch, int64_t n_out, int cin, int cout, int k3) {
    using A = typename AccOf<T>::type;
    constexpr int TX = TN / 4, TY = SIMT_THREADS / TX, TM = TY * 4;
    __shared__ A s_a[SIMT_KC][TM + 4];
    __shared__ A s_b[SIMT_KC][TN];
    __shared__ int s_idx[TM];

    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int64_t row0 = int64_t(blockIdx.x) * TM;
    const int col0 = blockIdx.y * TN;

    A acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j] = A(0);

    for (int k = 0; k < k3; ++k) {
        int any = 0;
        for (int r = tid; r < TM; r += SIMT_THREADS) {
            const int idx = row0 + r < n_out ? __ldg(nbr + int64_t(k) * pitch + row0 + r) : -1;
            s_idx[r] = idx;
            any |= idx >= 0;
        }
        if (!__syncthreads_or(any))
            continue; // no output row of this tile has a neighbour through tap k
        const T *wk = w + int64_t(k) * cin * cout;
        for (int c0 = 0; c0 < cin; c0 += SIMT_KC) {
            for (int e = tid; e < TM * SIMT_KC; e += SIMT_THREADS) {
                const int r = e / SIMT_KC, c = e % SIMT_KC;
                const int idx = s_idx[r];
                s_a[c][r] = (idx >= 0 && c0 + c < cin) ? to_acc<A>(x[int64_t(idx) * cin + c0 + c]) : A(0);
            }
            for (int e = tid; e < SIMT_KC * TN; e += SIMT_THREADS) {
                const int kk = e / TN, n = e % TN;
                s_b[kk][n] = (c0 + kk < cin && col0 + n < cout) ? to_acc<A>(wk[int64_t(c0 + kk) * cout + col0 + n]) : A(0);
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < SIMT_KC; ++kk) {
                A a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    a[i] = s_a[kk][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    b[j] = s_b[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[i][j] += a[i] * b[j];
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = row0 + ty * 4 + i;
        if (row >= n_out)
            continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = col0 + tx * 4 + j;
            if (col < cout) {
                A v = acc[i][j];
                if (bias)
                    v += to_acc<A>(bias[col]);
                y[row * cout + col] = from_acc<T>(v);
            }
        }
    }
}

template <typename T> static int launch_forward(const ConvArgs &a) {
    const T *x = reinterpret_cast<const T *>(a.x), *w = reinterpret_cast<const T *>(a.w),
            *bias = reinterpret_cast<const T *>(a.epi.bias);
    T *y = reinterpret_cast<T *>(a.y);
    if (a.cout <= 16) {
        constexpr int TN = 16, TM = SIMT_THREADS / (TN / 4) * 4;
        dim3 grid((unsigned)ceil_div(a.n_out, TM), 1);
        conv_os_simt_kernel<T, TN><<<grid, SIMT_THREADS, 0, a.stream>>>(x, w, bias, y, a.nbr, a.pitch, a.n_out, a.cin, a.cout, a.k3);
    } else if (a.cout <= 32) {
        constexpr int TN = 32, TM = SIMT_THREADS / (TN / 4) * 4;
        dim3 grid((unsigned)ceil_div(a.n_out, TM), 1);
        conv_os_simt_kernel<T, TN><<<grid, SIMT_THREADS, 0, a.stream>>>(x, w, bias, y, a.nbr, a.pitch, a.n_out, a.cin, a.cout, a.k3);
    } else {
        constexpr int TN = 64, TM = SIMT_THREADS / (TN / 4) * 4;
        dim3 grid((unsigned)ceil_div(a.n_out, TM), (unsigned)ceil_div(a.cout, TN));
        conv_os_simt_kernel<T, TN><<<grid, SIMT_THREADS, 0, a.stream>>>(x, w, bias, y, a.nbr, a.pitch, a.n_out, a.cin, a.cout, a.k3);
    }
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int simt_forward(const ConvArgs &a) {
    switch (a.dtype) {
    case FVC_F16: return launch_forward<__half>(a);
    case FVC_BF16: return launch_forward<__nv_bfloat16>(a);
    case FVC_F32: return launch_forward<float>(a);
    case FVC_F64: return launch_forward<double>(a);
    default: return set_error(FVC_ERR_UNSUPPORTED, "no convolution kernel for dtype code %d", a.dtype);
    }
}

// ---- weight gradient ------------------------------------------------------------------------------
// grid = (pair chunks, taps, Cin tiles * Cout tiles); a CTA reduces its chunk of tap k's pairs into a
// 64 x 64 tile of dW[k] and stores it to partial[chunk][k]; wgrad_reduce_kernel then sums the chunks in
// a fixed order and writes the public [Cout][Cin][K^3] layout.
template <typename T>
__global__ void __launch_bounds__(SIMT_THREADS)
wgrad_csr_simt_kernel(const T *__restrict__ x, const T *__restrict__ dy, const int32_t *__restrict__ gather,
                      const int32_t *__restrict__ scatter, const int64_t *__restrict__ offsets, int cin, int cout, int k3,
                      int64_t chunk, typename AccOf<T>::type *__restrict__ partial) {
    using A = typename AccOf<T>::type;
    constexpr int TILE = 64;
    __shared__ A s_a[SIMT_KC][TILE];
    __shared__ A s_b[SIMT_KC][TILE];
    __shared__ int s_g[SIMT_KC], s_s[SIMT_KC];

    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int k = blockIdx.y;
    const int co_tiles = (cout + TILE - 1) / TILE;
    const int ci0 = (blockIdx.z / co_tiles) * TILE, co0 = (blockIdx.z % co_tiles) * TILE;
    const int64_t seg_begin = offsets[k], seg_end = offsets[k + 1];
    const int64_t p_begin = seg_begin + int64_t(blockIdx.x) * chunk;
    const int64_t p_end = p_begin + chunk < seg_end ? p_begin + chunk : seg_end;

    A acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j] = A(0);

    for (int64_t p0 = p_begin; p0 < p_end; p0 += SIMT_KC) {
        if (tid < SIMT_KC) {
            const bool ok = p0 + tid < p_end;
            s_g[tid] = ok ? gather[p0 + tid] : -1;
            s_s[tid] = ok ? scatter[p0 + tid] : -1;
        }
        __syncthreads();
        for (int e = tid; e < SIMT_KC * TILE; e += SIMT_THREADS) {
            const int p = e / TILE, c = e % TILE;
            const int g = s_g[p], s = s_s[p];
            s_a[p][c] = (g >= 0 && ci0 + c < cin) ? to_acc<A>(x[int64_t(g) * cin + ci0 + c]) : A(0);
            s_b[p][c] = (s >= 0 && co0 + c < cout) ? to_acc<A>(dy[int64_t(s) * cout + co0 + c]) : A(0);
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < SIMT_KC; ++p) {
            A a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                a[i] = s_a[p][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = s_b[p][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
    A *dst = partial + (int64_t(blockIdx.x) * k3 + k) * cin * cout;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ci = ci0 + ty * 4 + i;
        if (ci >= cin)
            continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tx * 4 + j;
            if (co < cout)
                dst[int64_t(ci) * cout + co] = acc[i][j];
        }
    }
}

template <typename T>
__global__ void wgrad_reduce_kernel(const typename AccOf<T>::type *__restrict__ partial, int nchunks, int cin, int cout,
                                    int k3, T *__restrict__ grad_w) {
    using A = typename AccOf<T>::type;
    const int64_t per_chunk = int64_t(k3) * cin * cout;
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < per_chunk; e += int64_t(gridDim.x) * blockDim.x) {
        // e enumerates partial's [k][ci][co] layout (coalesced reads)
        const int k = int(e / (int64_t(cin) * cout));
        const int ci = int((e / cout) % cin), co = int(e % cout);
        A sum = A(0);
        for (int c = 0; c < nchunks; ++c)
            sum += partial[c * per_chunk + e];
        grad_w[(int64_t(co) * cin + ci) * k3 + k] = from_acc<T>(sum);
    }
}

static void chunking_for(int64_t max_n, int64_t *chunk, int64_t *nchunks);
static void wgrad_chunking(const int64_t *offsets_host, int64_t k3, int64_t *chunk, int64_t *nchunks) {
    int64_t max_n = 0;
    for (int64_t k = 0; k < k3; ++k)
        max_n = offsets_host[k + 1] - offsets_host[k] > max_n ? offsets_host[k + 1] - offsets_host[k] : max_n;
    chunking_for(max_n, chunk, nchunks);
}
static void chunking_for(int64_t max_n, int64_t *chunk, int64_t *nchunks) {
    int64_t c = ceil_div(max_n > 0 ? max_n : 1, 64);
    c = c < 2048 ? 2048 : c;
    c = ceil_div(c, SIMT_KC) * SIMT_KC;
    *chunk = c;
    *nchunks = ceil_div(max_n > 0 ? max_n : 1, c);
}

size_t simt_wgrad_scratch_bytes(int64_t max_pairs_per_tap, int32_t cin, int32_t cout, int64_t k3, int32_t dtype) {
    const size_t acc = dtype == FVC_F64 ? 8 : 4;
    int64_t chunk = 0, nchunks = 0;
    chunking_for(max_pairs_per_tap, &chunk, &nchunks); // the very split launch_wgrad uses (<= 64 chunks)
    return size_t(nchunks) * size_t(k3) * size_t(cin) * size_t(cout) * acc + 256;
}

template <typename T> static int launch_wgrad(const WgradArgs &a) {
    using A = typename AccOf<T>::type;
    int64_t chunk = 0, nchunks = 0;
    wgrad_chunking(a.offsets_host, a.k3, &chunk, &nchunks);
    const size_t need = size_t(nchunks) * size_t(a.k3) * size_t(a.cin) * size_t(a.cout) * sizeof(A);
    FVC_REQUIRE(a.scratch && a.scratch_bytes >= need, FVC_ERR_RUNTIME, "wgrad scratch too small: %zu < %zu", a.scratch_bytes,
                need);
    A *partial = reinterpret_cast<A *>(a.scratch);
    const int tiles = int(ceil_div(a.cin, 64) * ceil_div(a.cout, 64));
    FVC_REQUIRE(a.k3 <= 65535 && tiles <= 65535, FVC_ERR_UNSUPPORTED, "kernel volume / channel tiling exceeds grid limits");
    dim3 grid((unsigned)nchunks, (unsigned)a.k3, (unsigned)tiles);
    wgrad_csr_simt_kernel<T><<<grid, SIMT_THREADS, 0, a.stream>>>(
        reinterpret_cast<const T *>(a.x), reinterpret_cast<const T *>(a.dy), a.gather, a.scatter, a.offsets_dev, a.cin, a.cout,
        a.k3, chunk, partial);
    FVC_LAUNCH_CHECK();
    return wgrad_reduce_partials(partial, int(nchunks), a.cin, a.cout, a.k3, a.dtype, a.grad_w, a.stream);
}

template <typename T>
static int launch_reduce(const void *partial, int nchunks, int32_t cin, int32_t cout, int32_t k3, void *grad_w, cudaStream_t stream) {
    using A = typename AccOf<T>::type;
    const int64_t per_chunk = int64_t(k3) * cin * cout;
    const int blocks = int(ceil_div(per_chunk, 256) > 148 * 8 ? 148 * 8 : ceil_div(per_chunk, 256));
    wgrad_reduce_kernel<T><<<blocks, 256, 0, stream>>>(reinterpret_cast<const A *>(partial), nchunks, cin, cout, k3,
                                                       reinterpret_cast<T *>(grad_w));
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}

int wgrad_reduce_partials(const void *partial, int nchunks, int32_t cin, int32_t cout, int32_t k3, int32_t dtype, void *grad_w,
                          cudaStream_t stream) {
    switch (dtype) {
    case FVC_F16: return launch_reduce<__half>(partial, nchunks, cin, cout, k3, grad_w, stream);
    case FVC_BF16: return launch_reduce<__nv_bfloat16>(partial, nchunks, cin, cout, k3, grad_w, stream);
    case FVC_F32: return launch_reduce<float>(partial, nchunks, cin, cout, k3, grad_w, stream);
    case FVC_F64: return launch_reduce<double>(partial, nchunks, cin, cout, k3, grad_w, stream);
    default: return set_error(FVC_ERR_UNSUPPORTED, "no weight-gradient reduction for dtype code %d", dtype);
    }
}

int simt_wgrad(const WgradArgs &a) {
    switch (a.dtype) {
    case FVC_F16: return launch_wgrad<__half>(a);
    case FVC_BF16: return launch_wgrad<__nv_bfloat16>(a);
    case FVC_F32: return launch_wgrad<float>(a);
    case FVC_F64: return launch_wgrad<double>(a);
    default: return set_error(FVC_ERR_UNSUPPORTED, "no weight-gradient kernel for dtype code %d", a.dtype);
    }
}

} // namespace fvc
