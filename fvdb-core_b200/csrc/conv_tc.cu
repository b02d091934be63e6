// conv_tc.cu -- tcgen05 / TMEM implicit-GEMM path (placeholder while the CUDA-core path is validated).
#include "conv_internal.cuh"

namespace fvc {
bool tc_forward_supported(int32_t, int32_t, int64_t, int32_t) { return false; }
size_t tc_forward_scratch_bytes(int64_t, int32_t, int32_t, int64_t, int32_t) { return 0; }
int tc_forward(const ConvArgs &) { return set_error(FVC_ERR_UNSUPPORTED, "tensor-core path not built"); }
bool tc_wgrad_supported(int32_t, int32_t, int64_t, int32_t) { return false; }
size_t tc_wgrad_scratch_bytes(int64_t, int32_t, int32_t, int64_t, int32_t) { return 0; }
int tc_wgrad(const WgradArgs &) { return set_error(FVC_ERR_UNSUPPORTED, "tensor-core path not built"); }
} // namespace fvc
