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

extern "C" {

size_t fvc_conv_scratch_bytes(int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype) {
    return tc_forward_supported(cin, cout, kernel_volume, dtype) ? tc_forward_scratch_bytes(n_in, n_out, cin, cout, kernel_volume, dtype) : 0;
}

int fvc_conv_forward(const void *x, const void *w_packed, const void *bias, void *y, const int32_t *nbr, int64_t pitch,
                     const uint64_t *tile_mask, int64_t n_in, int64_t n_out, int32_t cin, int32_t cout, int64_t kernel_volume, int32_t dtype,
                     int32_t path, void *scratch, size_t scratch_bytes, fvc_stream_t stream) {
    int rc = check_common(x, w_packed, n_in, n_out, cin, cout, kernel_volume, dtype, "fvc_conv_forward");
    if (rc)
        return rc;
    FVC_REQUIRE(path >= 0 && path <= 2, FVC_ERR_VALUE, "path must be 0 (auto), 1 (CUDA-core) or 2 (tensor-core)");
    if (n_out == 0)
        return FVC_OK;
    FVC_REQUIRE(y && (kernel_volume == 0 || nbr), FVC_ERR_RUNTIME, "fvc_conv_forward: null output / map pointer");
    FVC_REQUIRE(pitch >= n_out, FVC_ERR_RUNTIME, "fvc_conv_forward: map pitch %lld < output rows %lld", (long long)pitch,
                (long long)n_out);
    ConvArgs a{x, w_packed, bias, y, nbr, pitch, tile_mask, n_in, n_out, cin, cout, int32_t(kernel_volume), dtype, scratch, scratch_bytes,
               reinterpret_cast<cudaStream_t>(stream)};
    const bool tc_ok = tc_forward_supported(cin, cout, kernel_volume, dtype);
    if (path == 2 && !tc_ok)
        return set_error(FVC_ERR_UNSUPPORTED, "tensor-core path does not admit dtype code %d with channels %d -> %d", dtype, cin, cout);
    if (tc_ok && path != 1)
        return tc_forward(a);
    return simt_forward(a);
}

size_t fvc_conv_wgrad_scratch_bytes(int64_t n_in, int64_t n_out, int64_t total_pairs, int32_t cin, int32_t cout, int64_t kernel_volume,
                                    int32_t dtype) {
    size_t simt = simt_wgrad_scratch_bytes(total_pairs, cin, cout, kernel_volume, dtype);
    size_t tc = tc_wgrad_supported(cin, cout, kernel_volume, dtype) ? tc_wgrad_scratch_bytes(n_in, n_out, cin, cout, kernel_volume, dtype) : 0;
    return simt > tc ? simt : tc;
}

int fvc_conv_wgrad(const void *x, const void *dy, const int32_t *gather, const int32_t *scatter,
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
    FVC_REQUIRE(offsets_host, FVC_ERR_RUNTIME, "fvc_conv_wgrad: offsets_host must be provided");
    if (n_out == 0 || n_in == 0 || offsets_host[kernel_volume] == 0) { // GatherScatterDefault.cu:771-777
        FVC_CUDA(cudaMemsetAsync(grad_w, 0, out_bytes, stream));
        return FVC_OK;
    }
    FVC_REQUIRE(gather && scatter && offsets_dev, FVC_ERR_RUNTIME, "fvc_conv_wgrad: null CSR pointer");
    WgradArgs a{x, dy, gather, scatter, offsets_host, offsets_dev, nbr, pitch, tile_mask, n_in, n_out, cin, cout, int32_t(kernel_volume), dtype,
                grad_w, scratch, scratch_bytes, stream};
    const bool tc_ok = nbr && tc_wgrad_supported(cin, cout, kernel_volume, dtype);
    if (path == 2 && !tc_ok)
        return set_error(FVC_ERR_UNSUPPORTED, "tensor-core wgrad does not admit dtype code %d with channels %d -> %d", dtype, cin, cout);
    if (tc_ok && path != 1)
        return tc_wgrad(a);
    return simt_wgrad(a);
}

} // extern "C"
