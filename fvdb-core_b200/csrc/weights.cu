// weights.cu -- weight re-layout: public [Cout,Cin,k0,k1,k2] (arbitrary strides) -> [K^3][Cin][Cout] or
// [K^3][Cout][Cin] in the compute dtype.  Replaces the per-call
// `weights.permute({2,3,4,1,0}).reshape({K,Cin,Cout}).contiguous()` + cast of
// GatherScatterDefault.cu:691-694,762-765.
#include "fvc_common.cuh"

namespace fvc {

__device__ __forceinline__ double load_any(const void *p, int64_t i, int dtype) {
    switch (dtype) {
    case FVC_F16: return double(__half2float(reinterpret_cast<const __half *>(p)[i]));
    case FVC_BF16: return double(__bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p)[i]));
    case FVC_F32: return double(reinterpret_cast<const float *>(p)[i]);
    default: return reinterpret_cast<const double *>(p)[i];
    }
}

__device__ __forceinline__ void store_any(void *p, int64_t i, int dtype, double v) {
    switch (dtype) {
    case FVC_F16: reinterpret_cast<__half *>(p)[i] = __float2half_rn(float(v)); break;
    case FVC_BF16: reinterpret_cast<__nv_bfloat16 *>(p)[i] = __float2bfloat16_rn(float(v)); break;
    case FVC_F32: reinterpret_cast<float *>(p)[i] = float(v); break;
    default: reinterpret_cast<double *>(p)[i] = v; break;
    }
}

struct Strides5 {
    int64_t s[5];
};

__global__ void pack_weights_kernel(const void *__restrict__ w, Strides5 st, int dtype_in, int cout, int cin, int k0,
                                    int k1, int k2, int layout, int flip, int dtype_out, void *__restrict__ out) {
    const int64_t k3 = int64_t(k0) * k1 * k2, total = k3 * cin * cout;
    for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        // e enumerates the OUTPUT layout so that stores coalesce
        int64_t k = e / (int64_t(cin) * cout);
        const int64_t rem = e - k * int64_t(cin) * cout;
        int ci, co;
        if (layout == 0) {
            ci = int(rem / cout), co = int(rem % cout);
        } else {
            co = int(rem / cin), ci = int(rem % cin);
        }
        const int64_t ksrc = flip ? k3 - 1 - k : k;
        const int t0 = int(ksrc / (int64_t(k1) * k2)), t1 = int((ksrc / k2) % k1), t2 = int(ksrc % k2);
        const int64_t src = co * st.s[0] + ci * st.s[1] + t0 * st.s[2] + t1 * st.s[3] + t2 * st.s[4];
        store_any(out, e, dtype_out, load_any(w, src, dtype_in));
    }
}

} // namespace fvc

using namespace fvc;

extern "C" int fvc_pack_weights(const void *weights, const int64_t strides[5], int32_t dtype_in, int32_t cout, int32_t cin,
                                int32_t k0, int32_t k1, int32_t k2, int32_t layout, int32_t flip_taps, int32_t dtype_out,
                                void *out, fvc_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    FVC_REQUIRE(dtype_size(dtype_in) && dtype_size(dtype_out), FVC_ERR_UNSUPPORTED, "unsupported weight dtype");
    FVC_REQUIRE(layout == 0 || layout == 1, FVC_ERR_VALUE, "layout must be 0 or 1");
    const int64_t total = int64_t(k0) * k1 * k2 * cin * cout;
    if (total == 0)
        return FVC_OK;
    Strides5 st;
    for (int d = 0; d < 5; ++d)
        st.s[d] = strides[d];
    const int blocks = int(ceil_div(total, 256) > 148 * 8 ? 148 * 8 : ceil_div(total, 256));
    pack_weights_kernel<<<blocks, 256, 0, stream>>>(weights, st, dtype_in, cout, cin, k0, k1, k2, layout, flip_taps,
                                                    dtype_out, out);
    FVC_LAUNCH_CHECK();
    return FVC_OK;
}
