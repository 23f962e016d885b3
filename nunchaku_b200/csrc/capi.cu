// C-ABI glue shared by all translation units: error string, launch counter, device check.
#include "common.cuh"

namespace nb200 {

namespace {
thread_local std::string g_last_error;
thread_local int g_launches = 0;
}  // namespace

void set_last_error(const std::string &msg) { g_last_error = msg; }

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

}  // namespace nb200

extern "C" __attribute__((visibility("default"))) int nb200_abi_version(void) { return NB200_ABI_VERSION; }

extern "C" __attribute__((visibility("default"))) const char *nb200_last_error(void) { return nb200::g_last_error.c_str(); }

extern "C" __attribute__((visibility("default"))) int nb200_last_launch_count(void) { return nb200::g_launches; }

extern "C" __attribute__((visibility("default"))) int nb200_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return nb200::fail(NB200_ERR_CUDA, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10)
        return nb200::fail(NB200_ERR_ARCH, "nunchaku_b200 kernels are sm_100a only; device is sm_" +
                                               std::to_string(major) + std::to_string(minor));
    return NB200_OK;
}
