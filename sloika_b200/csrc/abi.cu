// Library-level entry points of the C ABI (include/sloika_b200.h).
#include "common.cuh"

extern "C" int sloika_b200_abi_version(void) { return SLOIKA_B200_ABI_VERSION; }

extern "C" const char *sloika_b200_strerror(int code)
{
    switch (code) {
        case SLOIKA_OK:              return "success";
        case SLOIKA_ERR_ARG:         return "invalid argument (null pointer, bad size or inconsistent shape)";
        case SLOIKA_ERR_UNSUPPORTED: return "request not supported by the sm_100a kernels";
        case SLOIKA_ERR_WORKSPACE:   return "workspace missing or too small";
        default:
            if (code > 0) return cudaGetErrorString((cudaError_t)code);
            return "unknown sloika_b200 error";
    }
}

extern "C" int sloika_b200_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return (int)err;
    cudaDeviceProp prop;
    err = cudaGetDeviceProperties(&prop, dev);
    if (err != cudaSuccess) return (int)err;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return SLOIKA_OK;
}

namespace sloika { namespace gemm_tc { extern int sm_budget; } }

extern "C" int sloika_b200_set_gemm_sm_budget(int sms)
{
    if (sms < 0) return SLOIKA_ERR_ARG;
    sloika::gemm_tc::sm_budget = sms;
    return SLOIKA_OK;
}
