// tc_math.cuh -- small device helpers shared by the tensor-core executors' epilogues.
#pragma once
#include "fvc_common.cuh"

namespace fvc {

__device__ __forceinline__ uint32_t pack_half2(float a, float b, bool bf16) {
    if (bf16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float half_to_float(uint16_t v, bool bf16) {
    return bf16 ? __bfloat162float(*reinterpret_cast<__nv_bfloat16 *>(&v)) : __half2float(*reinterpret_cast<__half *>(&v));
}
__device__ __forceinline__ void unpack_half2(uint32_t p, bool bf16, float &a, float &b) {
    if (bf16) {
        a = __uint_as_float(p << 16), b = __uint_as_float(p & 0xFFFF0000u);
    } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&p));
        a = f.x, b = f.y;
    }
}

// Column sums across the 32 lanes of a warp by recursive halving: on return lane l holds the sum over all lanes of
// column (l & (N - 1)) in v[0]  (N - 1 (+1) shuffles instead of 5 N).
template <int N> __device__ __forceinline__ float warp_column_sum(float (&v)[N], int lane) {
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = upper ? v[i] : v[i + h], keep = upper ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    if (N < 32)
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
    return v[0];
}

} // namespace fvc
