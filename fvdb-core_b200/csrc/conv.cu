// conv.cu -- C-ABI execution entry points: precondition checks (GatherScatterDefault.cu:635-667) and the
// choice between the two hand-written kernel families.  This is a switch on dtype / channel count, not a
// multi-backend dispatch table (reference: dispatch_table.select at GatherScatterDefault.cu:854-858).
#include "conv_internal.cuh"

using namespace fvc;

static int check_common(const void *x, const void *w, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout,
                        int64_t k3, int32_t dtype, const char *name) {
    FVC_REQUIRE(dtype_size(dtype) != 0, FVC_ERR_UNSUPPORTED, "%s: features must be floating point (dtype code %d)", name, dtype);
    FVC_REQUIRE(cin > 0 && cout > 0, FVC_ERR_RUNTIME, "%s: channel counts must be positive, got %d -> %d", name, cin, cout);
    FVC_REQUIRE(n_in >= 0 && n_in <= INT32_MAX && n_out >= 0 && n_out <= INT32_MAX, FVC_ERR_RUNTIME,
                "%s: voxel counts exceed the int32 index limit", name);
    FVC_REQUIRE(k3 >= 0 && k3 <= INT32_MAX, FVC_ERR_RUNTIME, "%s: kernel volume out of range", name);
    FVC_REQUIRE((n_in == 0 || x) && (k3 == 0 || w), FVC_ERR_RUNTIME, "%s: null tensor pointer", name);
    return FVC_OK;
}

static int run_forward(const void *x, const void *w, bool w_prepared, bool x_split, const Epilogue &epi, void *y, const int32_t *nbr, int64_t pitch,
                       const uint64_t *tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype,
                       int32_t path, void *scratch, size_t scratch_bytes, fvc_stream_t stream) {
    int rc = check_common(x, w, n_in, n_out, cin, cout, kernel_volume, dtype, "fvc_conv_forward");
    if (rc)
        return rc;
    FVC_REQUIRE(path >= 0 && path <= 2, FVC_ERR_VALUE, "path must be 0 (auto), 1 (CUDA-core) or 2 (tensor-core)");
    if (n_out == 0)
        return FVC_OK;
    FVC_REQUIRE(y && (kernel_volume == 0 || nbr), FVC_ERR_RUNTIME, "fvc_conv_forward: null output / map pointer");
    FVC_REQUIRE(pitch >= n_out, FVC_ERR_RUNTIME, "fvc_conv_forward: map pitch %lld < output rows %lld", (long long)pitch,
                (long long)n_out);
    ConvArgs a{x, w, epi, y, nbr, pitch, tile_mask, n_in, n_out, cin, cout, int32_t(kernel_volume), dtype, scratch, scratch_bytes,
               reinterpret_cast<cudaStream_t>(stream), w_prepared, x_split};
    const bool tc_ok = tc_forward_supported(cin, cout, kernel_volume, dtype);
    if (path == 2 && !tc_ok)
        return set_error(FVC_ERR_UNSUPPORTED, "tensor-core path does not admit dtype code %d with channels %d -> %d", dtype, cin, cout);
    if (tc_ok && path != 1)
        return tc_forward(a);
    FVC_REQUIRE(!x_split, FVC_ERR_UNSUPPORTED, "fvc_conv_forward: split feature rows are a tensor-core (fp32) operand");
    FVC_REQUIRE(!epi.scale && !epi.shift && !epi.residual && !epi.relu && !epi.stats, FVC_ERR_UNSUPPORTED,
                "fvc_conv_forward_ex: the fused block epilogue runs on the tensor-core path only (dtype code %d, channels %d -> %d)", dtype, cin, cout);
    return simt_forward(a);
}

extern "C" {

int32_t fvc_conv_kernel_family(int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype, int32_t path, int32_t pass) {
    if (path == 1)
        return 1;
    if (pass == 2) // fused backward (cin / cout are the PUBLIC weight dimensions)
        return tc_bwd_fused_supported(cin, cout, kernel_volume, dtype) && tc_forward_supported(cout, cin, kernel_volume, dtype) ? 2 : 1;
    const bool tc = pass == 0 ? tc_forward_supported(cin, cout, kernel_volume, dtype) : tc_wgrad_supported(cin, cout, kernel_volume, dtype);
    return tc ? 2 : 1;
}

size_t fvc_conv_backward_fused_scratch_bytes(int64_t n_in, int32_t cin, int32_t cout, int64_t kernel_volume) {
    return tc_bwd_fused_scratch_bytes(n_in, cin, cout, kernel_volume);
}

int fvc_conv_backward_fused(const void *grad_output, const void *features, const void *w_prepared_transposed, const int32_t *in_map, int64_t pitch,
                            const uint64_t *in_tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume,
                            int32_t dtype, int32_t flip_taps, void *grad_features, void *grad_w, void *scratch, size_t scratch_bytes,
                            fvc_stream_t stream) {
    int rc = check_common(features, w_prepared_transposed, n_in, n_out, cin, cout, kernel_volume, dtype, "fvc_conv_backward_fused");
    if (rc)
        return rc;
    FVC_REQUIRE(n_in > 0 && n_out > 0 && kernel_volume > 0, FVC_ERR_RUNTIME, "fvc_conv_backward_fused: empty problem (the caller zero-fills, GatherScatterDefault.cu:771-777)");
    FVC_REQUIRE(grad_output && in_map && grad_features && grad_w, FVC_ERR_RUNTIME, "fvc_conv_backward_fused: null pointer");
    BwdFusedArgs a{grad_output, features, w_prepared_transposed, in_map, pitch, in_tile_mask, n_in, n_out, cin, cout, int32_t(kernel_volume), dtype,
                   flip_taps ? 1 : 0, grad_features, grad_w, scratch, scratch_bytes, reinterpret_cast<cudaStream_t>(stream)};
    return tc_bwd_fused(a);
}

int fvc_set_tuning(int32_t key, int32_t value) {
    FVC_REQUIRE(key == 0 || key == 1, FVC_ERR_VALUE, "unknown tuning key %d", key);
    (key == 0 ? g_tc_variant : g_wgrad_variant) = value;
    return FVC_OK;
}

size_t fvc_conv_scratch_bytes(int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype) {
    return tc_forward_supported(cin, cout, kernel_volume, dtype) ? tc_forward_scratch_bytes(n_in, n_out, cin, cout, kernel_volume, dtype) : 0;
}

int fvc_conv_forward(const void *x, const void *w_packed, const void *bias, void *y, const int32_t *nbr, int64_t pitch,
                     const uint64_t *tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype,
                     int32_t path, void *scratch, size_t scratch_bytes, fvc_stream_t stream) {
    Epilogue epi;
    epi.bias = bias;
    return run_forward(x, w_packed, false, false, epi, y, nbr, pitch, tile_mask, n_in, n_out, cin, cout, kernel_volume, dtype, path, scratch,
                       scratch_bytes, stream);
}

static inline bool takes_tc(int32_t cin, int32_t cout, int64_t k3, int32_t dtype, int32_t path) {
    return path != 1 && tc_forward_supported(cin, cout, k3, dtype);
}

size_t fvc_conv_weights_bytes(int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype, int32_t path) {
    if (takes_tc(cin, cout, kernel_volume, dtype, path))
        return tc_weight_image_bytes(cin, cout, kernel_volume, dtype);
    return align_up(size_t(kernel_volume) * size_t(cin) * size_t(cout) * dtype_size(dtype), 256);
}

int fvc_conv_prepare_weights(const void *weights, const int64_t strides[5], int32_t dtype_in, int32_t cout, int32_t cin, int32_t k0, int32_t k1,
                             int32_t k2, int32_t transpose, int32_t flip_taps, int32_t dtype, int32_t path, void *prepared, size_t prepared_bytes,
                             fvc_stream_t stream) {
    FVC_REQUIRE(dtype_size(dtype_in) && dtype_size(dtype), FVC_ERR_UNSUPPORTED, "unsupported weight dtype");
    FVC_REQUIRE(cin > 0 && cout > 0 && k0 > 0 && k1 > 0 && k2 > 0, FVC_ERR_RUNTIME, "fvc_conv_prepare_weights: empty weight tensor");
    const int64_t k3 = int64_t(k0) * k1 * k2;
    const int32_t cin_e = transpose ? cout : cin, cout_e = transpose ? cin : cout; // the executor's reduction / output channels
    FVC_REQUIRE(prepared && prepared_bytes >= fvc_conv_weights_bytes(cin_e, cout_e, k3, dtype, path), FVC_ERR_RUNTIME,
                "fvc_conv_prepare_weights: output buffer too small");
    FVC_REQUIRE((reinterpret_cast<uintptr_t>(prepared) & 255) == 0, FVC_ERR_RUNTIME, "fvc_conv_prepare_weights: output must be 256-byte aligned");
    if (takes_tc(cin_e, cout_e, k3, dtype, path))
        return tc_prepare_weights(weights, strides, dtype_in, cout, cin, k0, k1, k2, transpose, flip_taps, dtype, prepared,
                                  reinterpret_cast<cudaStream_t>(stream));
    return fvc_pack_weights(weights, strides, dtype_in, cout, cin, k0, k1, k2, transpose ? 1 : 0, flip_taps, dtype, prepared, stream);
}

int64_t fvc_conv_stats_blocks(int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype, int32_t path, int32_t *rows_per_block) {
    if (!takes_tc(cin, cout, kernel_volume, dtype, path)) {
        if (rows_per_block)
            *rows_per_block = 0;
        return 0; // the fused statistics exist on the tensor-core path only
    }
    return tc_stats_blocks(n_out, cin, cout, kernel_volume, dtype, rows_per_block);
}

int fvc_conv_forward_ex(const void *x, int32_t x_is_split, const void *w_prepared, const FvcConvEpilogue *epilogue, void *y, const int32_t *nbr,
                        int64_t pitch, const uint64_t *tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume,
                        int32_t dtype, int32_t path, void *scratch, size_t scratch_bytes, fvc_stream_t stream) {
    Epilogue epi;
    if (epilogue) {
        epi.bias = epilogue->bias, epi.scale = epilogue->scale, epi.shift = epilogue->shift, epi.residual = epilogue->residual;
        epi.relu = epilogue->relu, epi.stats = epilogue->stats;
    }
    return run_forward(x, w_prepared, true, x_is_split != 0, epi, y, nbr, pitch, tile_mask, n_in, n_out, cin, cout, kernel_volume, dtype, path, scratch,
                       scratch_bytes, stream);
}

int fvc_split_rows(const float *x, int64_t n, int32_t channels, void *split_rows, fvc_stream_t stream) {
    FVC_REQUIRE(channels > 0 && channels % 8 == 0, FVC_ERR_UNSUPPORTED, "fvc_split_rows: channel count %d must be a multiple of 8", channels);
    FVC_REQUIRE(n == 0 || (x && split_rows), FVC_ERR_RUNTIME, "fvc_split_rows: null pointer");
    FVC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(split_rows) & 15) == 0, FVC_ERR_RUNTIME,
                "fvc_split_rows: pointers must be 16-byte aligned");
    return tc_split_rows(x, n, channels, reinterpret_cast<uint16_t *>(split_rows), reinterpret_cast<cudaStream_t>(stream));
}

size_t fvc_conv_wgrad_scratch_bytes(int64_t n_in, int64_t n_out, int64_t max_pairs_per_tap, int32_t cin, int32_t cout, int64_t kernel_volume,
                                    int32_t dtype, int32_t path, int32_t has_dense_map) {
    // sized for the kernel family that will run (fvc_conv_wgrad makes the same choice)
    if (path != 1 && has_dense_map && tc_wgrad_supported(cin, cout, kernel_volume, dtype))
        return tc_wgrad_scratch_bytes(n_in, n_out, cin, cout, kernel_volume, dtype);
    return simt_wgrad_scratch_bytes(max_pairs_per_tap, cin, cout, kernel_volume, dtype);
}

int fvc_conv_wgrad(const void *x, const void *dy, const int32_t *gather, const int32_t *scatter,
                   const int64_t *offsets_host, const int64_t *offsets_dev, const int32_t *nbr, int64_t pitch,
                   const uint64_t *tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype,
                   int32_t path, void *grad_w, void *scratch, size_t scratch_bytes, fvc_stream_t stream_) {
    return fvc_conv_wgrad_ex(x, 0, dy, 0, gather, scatter, offsets_host, offsets_dev, nbr, pitch, tile_mask, n_in, n_out, cin, cout, kernel_volume, dtype,
                             path, grad_w, scratch, scratch_bytes, stream_);
}

int fvc_conv_wgrad_ex(const void *x, int32_t x_is_split, const void *dy, int32_t dy_is_split, const int32_t *gather, const int32_t *scatter,
                      const int64_t *offsets_host, const int64_t *offsets_dev, const int32_t *nbr, int64_t pitch,
                      const uint64_t *tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype,
                      int32_t path, void *grad_w, void *scratch, size_t scratch_bytes, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    // (dy stands in for the weights argument of the shared check; it may be NULL only when there are no output rows)
    int rc = check_common(x, n_out == 0 ? reinterpret_cast<const void *>(1) : dy, n_in, n_out, cin, cout, kernel_volume, dtype, "fvc_conv_wgrad");
    if (rc)
        return rc;
    FVC_REQUIRE(path >= 0 && path <= 2, FVC_ERR_VALUE, "path must be 0 (auto), 1 (CUDA-core) or 2 (tensor-core)");
    const size_t out_bytes = size_t(kernel_volume) * size_t(cin) * size_t(cout) * dtype_size(dtype);
    if (out_bytes == 0)
        return FVC_OK;
    FVC_REQUIRE(grad_w, FVC_ERR_RUNTIME, "fvc_conv_wgrad: null grad_w pointer");
    const bool tc_ok = nbr && tc_wgrad_supported(cin, cout, kernel_volume, dtype);
    const bool run_tc = tc_ok && path != 1;
    FVC_REQUIRE(offsets_host || run_tc, FVC_ERR_RUNTIME, "fvc_conv_wgrad: offsets_host must be provided");
    if (n_out == 0 || n_in == 0 || (offsets_host && offsets_host[kernel_volume] == 0)) { // GatherScatterDefault.cu:771-777
        FVC_CUDA(cudaMemsetAsync(grad_w, 0, out_bytes, stream));
        return FVC_OK;
    }
    FVC_REQUIRE(run_tc || (gather && scatter && offsets_dev), FVC_ERR_RUNTIME, "fvc_conv_wgrad: null CSR pointer");
    WgradArgs a{x, dy, gather, scatter, offsets_host, offsets_dev, nbr, pitch, tile_mask, n_in, n_out, cin, cout, int32_t(kernel_volume), dtype,
                grad_w, scratch, scratch_bytes, stream, x_is_split != 0, dy_is_split != 0};
    if (path == 2 && !tc_ok)
        return set_error(FVC_ERR_UNSUPPORTED, "tensor-core wgrad does not admit dtype code %d with channels %d -> %d", dtype, cin, cout);
    if (tc_ok && path != 1)
        return tc_wgrad(a);
    FVC_REQUIRE(!a.x_split && !a.dy_split, FVC_ERR_UNSUPPORTED, "fvc_conv_wgrad_ex: split rows are a tensor-core (fp32) operand");
    return simt_wgrad(a);
}

} // extern "C"
