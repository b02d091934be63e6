// conv_internal.cuh -- internal entry points shared by conv.cu (dispatch), conv_simt.cu, conv_tc.cu.
#pragma once
#include "fvc_common.cuh"

namespace fvc {

// Fused epilogue of the forward / dgrad kernels (FvcConvEpilogue in the ABI); every member optional.
// stored = act2(act1((acc + bias) * scale + shift) + residual), act1 / act2 = ReLU when bit 0 / bit 1 of `relu` is set;
// stats = per-CTA column sums of the stored values.
struct Epilogue {
    const void *bias = nullptr;     // [Cout] in `dtype`
    const float *scale = nullptr;   // [Cout] fp32
    const float *shift = nullptr;   // [Cout] fp32
    const void *residual = nullptr; // [n_out][Cout] in `dtype`
    int relu = 0;
    float *stats = nullptr;         // [stats blocks][2][Cout] fp32 partial (sum, sum of squares) of the stored values
};

struct ConvArgs {
    const void *x;
    const void *w; // [K^3][Cin][Cout] in `dtype`, or (w_prepared) the executor's own weight image
    Epilogue epi;
    void *y;
    const int32_t *nbr; // tap-major dense map
    int64_t pitch;
    const uint64_t *tile_mask; // per 128-row tile tap bitmask (may be null)
    int64_t n_in, n_out;
    int32_t cin, cout;
    int32_t k3;
    int32_t dtype;
    void *scratch;
    size_t scratch_bytes;
    cudaStream_t stream;
    bool w_prepared = false; // w is the blob written by fvc_conv_prepare_weights for the path that runs
    bool x_split = false;    // fp32 only: x already holds the bf16 split rows [N][3][Cin] (fvc_split_rows)
};

struct WgradArgs {
    const void *x;
    const void *dy;
    const int32_t *gather, *scatter;
    const int64_t *offsets_host, *offsets_dev;
    const int32_t *nbr;
    int64_t pitch;
    const uint64_t *tile_mask;
    int64_t n_in, n_out;
    int32_t cin, cout;
    int32_t k3;
    int32_t dtype;
    void *grad_w; // [Cout][Cin][K^3] in `dtype`
    void *scratch;
    size_t scratch_bytes;
    cudaStream_t stream;
    bool x_split = false, dy_split = false; // fp32 only: operands already are bf16 split rows (fvc_split_rows)
};

// Fused backward (conv_tc_bwd.cu): dgrad and wgrad from one gather of grad_output, narrow channels only
struct BwdFusedArgs {
    const void *dy;     // grad_output rows [n_out][cout]
    const void *x;      // feature rows [n_in][cin]
    const void *w_img;  // fvc_conv_prepare_weights(transpose = 1, flip_taps) image for the tensor-core path
    const int32_t *map; // input-stationary dense map [K^3][pitch]: for input row i and tap k the output row, or -1
    int64_t pitch;
    const uint64_t *tile_mask; // its per-tile tap bitmask (may be null)
    int64_t n_in, n_out;
    int32_t cin, cout, k3, dtype;
    int32_t flip_taps; // the map is the forward map of a symmetric plan walked with mirrored taps: dW taps are stored mirrored
    void *dx;          // [n_in][cin]
    void *grad_w;      // [Cout][Cin][K^3] public layout
    void *scratch;
    size_t scratch_bytes;
    cudaStream_t stream;
};
bool tc_bwd_fused_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
size_t tc_bwd_fused_scratch_bytes(int64_t n_in, int32_t cin, int32_t cout, int64_t k3);
int tc_bwd_fused(const BwdFusedArgs &a);

// CUDA-core path: every dtype, every channel count
int simt_forward(const ConvArgs &a);
size_t simt_wgrad_scratch_bytes(int64_t total_pairs_max_tap, int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
int simt_wgrad(const WgradArgs &a);
// grad_w[co][ci][k] = sum over chunks of partial[chunk][k][ci][co] (fixed order), cast to `dtype`
int wgrad_reduce_partials(const void *partial, int nchunks, int32_t cin, int32_t cout, int32_t k3, int32_t dtype, void *grad_w,
                          cudaStream_t stream);

// tcgen05 path: f16/bf16 (and fp32 as a three-way bf16 split, forward/dgrad), channel counts the UMMA tile shapes admit
bool tc_forward_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
size_t tc_forward_scratch_bytes(int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
int tc_forward(const ConvArgs &a);
// weight image of the tensor-core executors, written straight from the public [Cout,Cin,k0,k1,k2] layout (any strides)
size_t tc_weight_image_bytes(int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
int tc_prepare_weights(const void *weights, const int64_t strides[5], int32_t dtype_in, int32_t cout, int32_t cin, int32_t k0, int32_t k1,
                       int32_t k2, int32_t transpose, int32_t flip_taps, int32_t dtype, void *image, cudaStream_t stream);
// rows of the per-CTA statistics partials the forward kernel writes for n_out output rows (Epilogue::stats), and rows per block
int64_t tc_stats_blocks(int64_t n_out, int32_t cin, int32_t cout, int64_t k3, int32_t dtype, int32_t *rows_per_block);
// narrow half-precision layers: gathered operand through tensor memory (conv_tc_ts.cu); same weight image as conv_tc.cu
bool tc_ts_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
int64_t tc_ts_stats_blocks(int64_t n_out, int32_t *rows_per_block);
int tc_ts_forward(const ConvArgs &a, const void *x, const uint8_t *img);
int tc_split_rows(const float *x, int64_t n, int c, uint16_t *xs, cudaStream_t stream);
// experiment knob (fvc_set_tuning): pipeline shape variant of the forward kernel, 0 = default
extern int g_tc_variant;
extern int g_wgrad_variant; // same for the weight-gradient kernel (key 1)
// opt in to > 48 KB of dynamic shared memory once per (kernel instantiation, device)
template <typename K> inline int ensure_dynamic_smem(K kernel, size_t bytes, std::atomic<unsigned long long> &done_mask) {
    int dev = 0;
    FVC_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (dev >= 64 || !(done_mask.load(std::memory_order_acquire) & bit)) {
        FVC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
        done_mask.fetch_or(bit, std::memory_order_release);
    }
    return FVC_OK;
}
bool tc_wgrad_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
size_t tc_wgrad_scratch_bytes(int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
int tc_wgrad(const WgradArgs &a);

} // namespace fvc
