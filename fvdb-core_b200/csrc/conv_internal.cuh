// conv_internal.cuh -- internal entry points shared by conv.cu (dispatch), conv_simt.cu, conv_tc.cu.
#pragma once
#include "fvc_common.cuh"

namespace fvc {

struct ConvArgs {
    const void *x;
    const void *w; // [K^3][Cin][Cout] in `dtype`
    const void *bias;
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
};

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
bool tc_wgrad_supported(int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
size_t tc_wgrad_scratch_bytes(int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t k3, int32_t dtype);
int tc_wgrad(const WgradArgs &a);

} // namespace fvc
