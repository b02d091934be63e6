// abi.cu -- error plumbing, device info and the host-only geometry entry points of libfvdbconv.
#include "fvc_common.cuh"

#include <limits>
#include <string>

namespace fvc {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launch_count{0};

int set_error(int code, const char *fmt, ...) {
    char buf[1024];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buf, sizeof(buf), fmt, args);
    va_end(args);
    t_last_error = buf;
    return code;
}

static int check_geometry(const int32_t kernel_size[3], const int32_t stride[3], int64_t *volume_out) {
    // ConvolutionGeometry.h:150-172 (strictly positive) and :186-198 (volume fits int64)
    int64_t volume = 1;
    for (int d = 0; d < 3; ++d) {
        FVC_REQUIRE(kernel_size[d] > 0, FVC_ERR_VALUE, "kernel_size must be strictly positive, got %d in dimension %d",
                    kernel_size[d], d);
        FVC_REQUIRE(stride[d] > 0, FVC_ERR_VALUE, "stride must be strictly positive, got %d in dimension %d", stride[d], d);
    }
    for (int d = 0; d < 3; ++d) {
        FVC_REQUIRE(volume <= std::numeric_limits<int64_t>::max() / kernel_size[d], FVC_ERR_VALUE,
                    "kernel volume overflows int64 for kernel_size [%d, %d, %d]", kernel_size[0], kernel_size[1],
                    kernel_size[2]);
        volume *= kernel_size[d];
    }
    if (volume_out)
        *volume_out = volume;
    return FVC_OK;
}

} // namespace fvc

using namespace fvc;

extern "C" {

int fvc_abi_version(void) { return FVC_ABI_VERSION; }

const char *fvc_last_error(void) { return t_last_error.c_str(); }

int64_t fvc_launch_count(void) { return g_launch_count.load(); }

int fvc_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    FVC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    FVC_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count)
        *sm_count = prop.multiProcessorCount;
    if (cc_major)
        *cc_major = prop.major;
    if (cc_minor)
        *cc_minor = prop.minor;
    return FVC_OK;
}

int fvc_geometry(const int32_t kernel_size[3], const int32_t stride[3], int32_t padding_before[3],
                 int32_t padding_after[3], int64_t *kernel_volume) {
    int rc = check_geometry(kernel_size, stride, kernel_volume);
    if (rc)
        return rc;
    for (int d = 0; d < 3; ++d) {
        padding_before[d] = (kernel_size[d] - 1) / 2;               // ConvolutionGeometry.h:174-177
        padding_after[d] = kernel_size[d] - 1 - padding_before[d]; // :180-183
    }
    return FVC_OK;
}

int fvc_geometry_tap_coord(const int32_t kernel_size[3], int64_t tap_index, int32_t tap[3]) {
    const int32_t one[3] = {1, 1, 1};
    int64_t volume = 0;
    int rc = check_geometry(kernel_size, one, &volume);
    if (rc)
        return rc;
    FVC_REQUIRE(tap_index >= 0 && tap_index < volume, FVC_ERR_INDEX, "tap index %lld out of range [0, %lld)",
                (long long)tap_index, (long long)volume);
    const int64_t yz = int64_t(kernel_size[1]) * kernel_size[2];
    tap[0] = int32_t(tap_index / yz);
    tap[1] = int32_t((tap_index / kernel_size[2]) % kernel_size[1]);
    tap[2] = int32_t(tap_index % kernel_size[2]);
    return FVC_OK;
}

int fvc_geometry_fine_from_coarse(const int32_t kernel_size[3], const int32_t stride[3], const int32_t coarse[3],
                                  const int32_t tap[3], int32_t fine[3]) {
    int rc = check_geometry(kernel_size, stride, nullptr);
    if (rc)
        return rc;
    for (int d = 0; d < 3; ++d)
        fine[d] = coarse[d] * stride[d] + tap[d] - (kernel_size[d] - 1) / 2;
    return FVC_OK;
}

int fvc_geometry_coarse_from_fine(const int32_t kernel_size[3], const int32_t stride[3], const int32_t fine[3],
                                  const int32_t tap[3], int32_t coarse[3], int32_t *divisible) {
    int rc = check_geometry(kernel_size, stride, nullptr);
    if (rc)
        return rc;
    int64_t q[3];
    for (int d = 0; d < 3; ++d) {
        const int64_t numer = int64_t(fine[d]) - (tap[d] - (kernel_size[d] - 1) / 2);
        int64_t r = numer % stride[d];
        if (r < 0)
            r += stride[d];
        if (r != 0) {
            *divisible = 0;
            return FVC_OK;
        }
        q[d] = (numer - r) / stride[d];
    }
    for (int d = 0; d < 3; ++d)
        coarse[d] = int32_t(q[d]);
    *divisible = 1;
    return FVC_OK;
}

} // extern "C"
