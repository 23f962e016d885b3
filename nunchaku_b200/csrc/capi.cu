// C-ABI glue shared by all translation units: error string, launch counter, device check.
#include <cuda.h>

#include <cstdlib>
#include <map>
#include <mutex>

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

int set_max_smem_once(const void *kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> done;
    int dev = 0;
    NB200_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find({kernel, dev});
    if (it != done.end() && it->second >= bytes) return NB200_OK;
    NB200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    done[{kernel, dev}] = bytes;
    return NB200_OK;
}

int current_device_sms(int *num_sms) {
    static std::mutex mu;
    static std::map<int, int> sms;
    int dev = 0;
    NB200_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto it = sms.find(dev);
    if (it == sms.end()) {
        int n = 0;
        NB200_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        it = sms.emplace(dev, n).first;
    }
    *num_sms = it->second;
    return NB200_OK;
}

bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("NB200_PDL");
        return e == nullptr || atoi(e) != 0;
    }();
    return on;
}
void reset_launch_count() { g_launches = 0; }

// ---- TMA tensor maps (driver entry point resolved through the runtime: no -lcuda) ------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

int make_map_2d(CUtensorMap *map, CUtensorMapDataType dt, const void *base, uint64_t inner, uint64_t rows,
                uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) return fail(NB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(NB200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
    return NB200_OK;
}


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
