// Shared helpers: dtype traits, the approximate-math primitives the reference uses (so that
// 4-bit codes land on the same side of rounding boundaries), error plumbing.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/nunchaku_b200.h"

namespace nb200 {

// ---- error plumbing (host) -----------------------------------------------------------------
void set_last_error(const std::string &msg);
int fail(int code, const std::string &msg);
void count_launch(int n = 1);
void reset_launch_count();
// 2-D row-major tensor map: `inner` elements per row, box = box_inner x box_rows
int make_map_2d(CUtensorMap *map, CUtensorMapDataType dt, const void *base, uint64_t inner, uint64_t rows,
                uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle swz);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device, and one process may
// drive several GPUs (ops run with the device of their tensors current).  Returns a status code.
int set_max_smem_once(const void *kernel, size_t bytes);
// SM count of the current device (cached per device)
int current_device_sms(int *num_sms);
// NB200_PDL=0 disables programmatic dependent launch (default on)
bool pdl_enabled();

// launch configuration with the optional cluster dimension and the programmatic-dependent-launch attribute
struct LaunchCfg {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t stream, unsigned cluster_x = 1) {
        cfg = cudaLaunchConfig_t{};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        unsigned n = 0;
        if (cluster_x > 1) {
            attr[n].id = cudaLaunchAttributeClusterDimension;
            attr[n].val.clusterDim.x = cluster_x;
            attr[n].val.clusterDim.y = 1;
            attr[n].val.clusterDim.z = 1;
            n++;
        }
        if (pdl_enabled()) {
            attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[n].val.programmaticStreamSerializationAllowed = 1;
            n++;
        }
        cfg.attrs = attr;
        cfg.numAttrs = n;
    }
};

#define NB200_CUDA_CHECK(expr)                                                                      \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            return ::nb200::fail(NB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
        }                                                                                           \
    } while (0)

#define NB200_REQUIRE(cond, msg)                                                        \
    do {                                                                                \
        if (!(cond)) return ::nb200::fail(NB200_ERR_INVALID_ARGUMENT, std::string(msg) + " [" #cond "]"); \
    } while (0)

// ---- 16-bit float traits ---------------------------------------------------------------------
template <typename T>
struct HalfTraits;

template <>
struct HalfTraits<__half> {
    using T = __half;
    using T2 = __half2;
    static constexpr bool kIsBf16 = false;
    __device__ __forceinline__ static float to_float(T v) { return __half2float(v); }
    __device__ __forceinline__ static T from_float(float v) { return __float2half_rn(v); }
    __device__ __forceinline__ static float2 to_float2(T2 v) { return __half22float2(v); }
    __device__ __forceinline__ static T2 from_float2(float2 v) { return __float22half2_rn(v); }
};

template <>
struct HalfTraits<__nv_bfloat16> {
    using T = __nv_bfloat16;
    using T2 = __nv_bfloat162;
    static constexpr bool kIsBf16 = true;
    __device__ __forceinline__ static float to_float(T v) { return __bfloat162float(v); }
    __device__ __forceinline__ static T from_float(float v) { return __float2bfloat16_rn(v); }
    __device__ __forceinline__ static float2 to_float2(T2 v) { return __bfloat1622float2(v); }
    __device__ __forceinline__ static T2 from_float2(float2 v) { return __float22bfloat162_rn(v); }
};

// ---- approximate math, same PTX as the reference (gemm_utils.cuh:247-344) ---------------------
__device__ __forceinline__ float rcp_approx_ftz(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx_ftz(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float tanh_approx(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// x * sigmoid(x), sigmoid = rcp(1 + 2^(-x*log2e))   (cuda_sigmoidf + silu)
__device__ __forceinline__ float silu_f32(float x) {
    const float e = ex2_approx_ftz(-1.442695041f * x);
    return x * rcp_approx_ftz(e + 1.0f);
}
// tanh-GELU exactly as gelu_half2
__device__ __forceinline__ float gelu_f32(float x) {
    const float x3 = x * x * x;
    const float t = 0.5f + 0.5f * tanh_approx(0.79788456f * (x + (0.044715f * x3)));
    return x * t;
}

// ---- 4-bit packing ----------------------------------------------------------------------------
// d = (c << 8) | (sat4(a) << 4) | sat4(b)
__device__ __forceinline__ uint32_t pack_sat_s4(int a, int b, uint32_t c) {
    uint32_t d;
    asm("cvt.pack.sat.s4.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t pack_sat_u4(int a, int b, uint32_t c) {
    uint32_t d;
    asm("cvt.pack.sat.u4.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int cvt_rni(float v) {
    int r;
    asm("cvt.rni.s32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
// byte = (e2m1(hi) << 4) | e2m1(lo)
__device__ __forceinline__ uint32_t cvt_e2m1x2(float hi, float lo) {
    uint32_t r;
    asm("{ .reg .b8 t; cvt.rn.satfinite.e2m1x2.f32 t, %1, %2; cvt.u32.u8 %0, t; }" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// 16 bits = (e4m3(hi) << 8) | e4m3(lo)
__device__ __forceinline__ uint32_t cvt_e4m3x2(float hi, float lo) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}

// Eight values (K order e0..e7) -> one u32 in the B200 INT4 layout: nibble p = e(2p), nibble
// p+4 = e(2p+1), signed values stored offset-binary (q + 8).
template <bool UNSIGNED>
__device__ __forceinline__ uint32_t pack8_int4_b200(const int (&q)[8]) {
    // byte0 = {e2,e0}  byte1 = {e6,e4}  byte2 = {e3,e1}  byte3 = {e7,e5}   (high nibble first)
    // (the c operand of cvt.pack is left 0 and the bytes are merged explicitly)
    uint32_t b0, b1, b2, b3;
    if constexpr (UNSIGNED) {
        b0 = pack_sat_u4(q[2], q[0], 0);
        b1 = pack_sat_u4(q[6], q[4], 0);
        b2 = pack_sat_u4(q[3], q[1], 0);
        b3 = pack_sat_u4(q[7], q[5], 0);
    } else {
        b0 = pack_sat_s4(q[2], q[0], 0);
        b1 = pack_sat_s4(q[6], q[4], 0);
        b2 = pack_sat_s4(q[3], q[1], 0);
        b3 = pack_sat_s4(q[7], q[5], 0);
    }
    uint32_t w = (b0 & 0xFFu) | ((b1 & 0xFFu) << 8) | ((b2 & 0xFFu) << 16) | ((b3 & 0xFFu) << 24);
    if constexpr (!UNSIGNED) w ^= 0x88888888u;
    return w;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_v4(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

}  // namespace nb200
