// AWQ W4A16 GEMV for the AdaLN modulation linears.
//
// Replaces gemv_awq / gemv_kernel (reference src/kernels/awq/gemv_awq.cu:101-294; caller GEMV_AWQ::forward, src/Linear.cpp:56-86;
// Python nunchaku/ops/gemv.py:10-58): out[m, n] = sum_k x[m, k] * (code[n, k] * scale[k/64, n] + zero[k/64, n]), M <= 8.
// HBM bound: the 4-bit weights (1.6 GB per FLUX step) are read exactly once, straight from the checkpoint layout -- no repack:
//
//   qweight int32 [OC/4, IC/8*4]: output channels in blocks of 8 = 2 groups of 4 interleaved rows; a group is stored as
//   [IC/64][row 4][64 k] codes, i.e. 128 contiguous bytes per 64-k chunk, and inside a run of 32 k (4 u32 w0..w3) element
//   8*ii + 2*jj + e is nibble ii + 4*e of w_jj -- so (w_jj >> 4*ii) & 0x000F000F is the adjacent pair (8ii + 2jj, 8ii + 2jj + 1) as two
//   16-bit lanes, which the magic-number trick turns into an hT2 of exact integers.
//
//   one warp = one group of 4 rows: lane l reads 16 bytes = 32 k of row (l % 8) / 2 at k = 256 it + 64 (l / 8) + 32 (l % 2), so a warp load is
//   512 contiguous bytes; 4 loads in flight per lane; x is staged once per CTA in shared memory (lanes of different rows broadcast);
//   arithmetic exactly as the reference (hT fma for the dequant, hT multiply, fp32 accumulate -- gemv_awq.cu:207-236), fixed-order
//   shuffle reduction over the 8 lanes of a row, one rounding to hT.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace nb200 {
namespace {

constexpr int kWarps = 8;            // 8 groups of 4 output channels per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kMaxM = 8;

// FUSE_SILU: x <- silu(x) while it is staged (the modulation linears see silu(temb): src/FluxModel.cpp AdaLayerNormZero::forward runs
// kernels::silu first) with the activation kernel's arithmetic (x / (1 + expf(-x)) in fp32, one rounding: activation_kernels_impl.cuh:7-10);
// bias: added to the rounded result in hT like the caller's `out += bias` -- both bit-identical to the separate launches they replace.
template <typename hT, int M, bool FUSE_SILU>
__global__ void __launch_bounds__(kThreads) gemv_awq_kernel(const hT *__restrict__ x, const uint4 *__restrict__ qw, const hT *__restrict__ scales,
                                                            const hT *__restrict__ zeros, const hT *__restrict__ bias, hT *__restrict__ out, int OC,
                                                            int IC) {
    using Tr = HalfTraits<hT>;
    using T2 = typename Tr::T2;
    extern __shared__ uint4 xs4[];   // [M][IC] hT
    ptx::griddep_launch_dependents();
    ptx::griddep_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < M * IC / 8; i += kThreads) {
        uint4 v = reinterpret_cast<const uint4 *>(x)[i];
        if constexpr (FUSE_SILU) {
            hT *e = reinterpret_cast<hT *>(&v);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float f = Tr::to_float(e[k]);
                e[k] = Tr::from_float(f / (1.0f + expf(-f)));
            }
        }
        xs4[i] = v;
    }
    __syncthreads();
    const int group = blockIdx.x * kWarps + warp;    // 4 output channels
    if (group * 4 >= OC) return;
    const int r = (lane & 7) >> 1, h = lane & 1, c = lane >> 3;
    const int oc = group * 4 + r;
    // group g of block b = g / 2, idx = g % 2: u32 offset b * IC + idx * 4 * IC / 8 = g * IC / 2; as uint4: g * IC / 8
    const uint4 *wp = qw + static_cast<size_t>(group) * (IC / 8) + lane;
    constexpr uint32_t kMagic = Tr::kIsBf16 ? 0x43004300u : 0x64006400u;    // 128 + code | 1024 + code, exact in hT
    T2 off2;
    off2.x = Tr::from_float(Tr::kIsBf16 ? 128.f : 1024.f);
    off2.y = off2.x;
    float acc[M];
#pragma unroll
    for (int m = 0; m < M; m++) acc[m] = 0.f;
    const int iters = IC / 256;
    constexpr int U = 4;
    for (int it0 = 0; it0 < iters; it0 += U) {
        uint4 w[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (it0 + u < iters) w[u] = ldg_nc_v4(wp + static_cast<size_t>(it0 + u) * 32);
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (it0 + u >= iters) break;
            const int k0 = (it0 + u) * 256 + c * 64 + h * 32;
            const int grp = k0 >> 6;
            T2 s2, z2;
            s2.x = scales[static_cast<size_t>(grp) * OC + oc];
            s2.y = s2.x;
            z2.x = zeros[static_cast<size_t>(grp) * OC + oc];
            z2.y = z2.x;
            const uint32_t words[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
            T2 wv[16];   // pair p = 4 ii + jj holds elements (8 ii + 2 jj, +1)
#pragma unroll
            for (int jj = 0; jj < 4; jj++)
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    uint32_t bits = ((words[jj] >> (4 * ii)) & 0x000F000Fu) | kMagic;
                    const T2 q = __hsub2(*reinterpret_cast<T2 *>(&bits), off2);      // exact integer 0..15
                    wv[4 * ii + jj] = __hfma2(q, s2, z2);                            // the reference's dequant (gemv_awq.cu:207-215)
                }
#pragma unroll
            for (int m = 0; m < M; m++) {
                const uint4 *xp = xs4 + (static_cast<size_t>(m) * IC + k0) / 8;
                float a = acc[m];
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    const uint4 xv = xp[v];
                    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 f = Tr::to_float2(__hmul2(wv[4 * v + e], *reinterpret_cast<const T2 *>(&xw[e])));   // hT product, fp32 sum
                        a += f.x;
                        a += f.y;
                    }
                }
                acc[m] = a;
            }
        }
    }
#pragma unroll
    for (int m = 0; m < M; m++) {
        float v = acc[m];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (h == 0 && c == 0) {
            hT y = Tr::from_float(v);
            if (bias != nullptr) y = Tr::from_float(Tr::to_float(y) + Tr::to_float(bias[oc]));
            out[static_cast<size_t>(m) * OC + oc] = y;
        }
    }
}

template <typename hT, bool FUSE_SILU>
int launch_gemv(const void *x, const void *qw, const void *scales, const void *zeros, const void *bias, void *out, int M, int OC, int IC,
                cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(M) * IC * sizeof(hT);
    const int grid = (OC / 4 + kWarps - 1) / kWarps;
#define NB200_GEMV_CASE(MM)                                                                                                                        \
    case MM: {                                                                                                                                     \
        auto kern = gemv_awq_kernel<hT, MM, FUSE_SILU>;                                                                                                       \
        if (smem > 48 * 1024)                                                                                                                      \
            if (int rc = set_max_smem_once(reinterpret_cast<const void *>(kern), smem)) return rc;                                                 \
        LaunchCfg lc(dim3(grid), dim3(kThreads), smem, stream);                                                                                    \
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, static_cast<const hT *>(x), static_cast<const uint4 *>(qw), static_cast<const hT *>(scales), \
                                            static_cast<const hT *>(zeros), static_cast<const hT *>(bias), static_cast<hT *>(out), OC, IC));       \
        break;                                                                                                                                     \
    }
    switch (M) {
        NB200_GEMV_CASE(1)
        NB200_GEMV_CASE(2)
        NB200_GEMV_CASE(3)
        NB200_GEMV_CASE(4)
        NB200_GEMV_CASE(5)
        NB200_GEMV_CASE(6)
        NB200_GEMV_CASE(7)
        NB200_GEMV_CASE(8)
        default: return fail(NB200_ERR_INVALID_ARGUMENT, "gemv_awq: M must be in 1..8");
    }
#undef NB200_GEMV_CASE
    count_launch();
    return NB200_OK;
}

}  // namespace
}  // namespace nb200

// bias (hT [OC]) may be NULL; fuse_silu != 0 applies SiLU to x first (the AdaLN modulation path: silu -> gemv -> + bias in one launch)
extern "C" __attribute__((visibility("default"))) int nb200_gemv_awq_fused(int dtype, const void *x, const void *qweight, const void *scales,
                                                                           const void *zeros, const void *bias, void *out, int M, int OC, int IC,
                                                                           int group_size, int fuse_silu, void *stream_) {
    using namespace nb200;
    reset_launch_count();
    NB200_REQUIRE(x && qweight && scales && zeros && out, "NULL tensor");
    NB200_REQUIRE(M >= 1 && M <= kMaxM, "M must be in 1..8 (gemv_awq.cu:276)");
    NB200_REQUIRE(group_size == 64, "group_size must be 64 (gemv_awq.cu:277)");
    NB200_REQUIRE(OC > 0 && OC % 8 == 0, "OC must be a positive multiple of 8");
    NB200_REQUIRE(IC > 0 && IC % 256 == 0, "IC must be a positive multiple of 256");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "dtype must be fp16 or bf16");
    NB200_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(qweight)) & 15) == 0, "x / qweight must be 16-byte aligned");
    NB200_REQUIRE(static_cast<size_t>(M) * IC * 2 <= 200 * 1024, "M * IC does not fit shared memory");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (dtype == NB200_BF16)
        return fuse_silu ? launch_gemv<__nv_bfloat16, true>(x, qweight, scales, zeros, bias, out, M, OC, IC, stream)
                         : launch_gemv<__nv_bfloat16, false>(x, qweight, scales, zeros, bias, out, M, OC, IC, stream);
    return fuse_silu ? launch_gemv<__half, true>(x, qweight, scales, zeros, bias, out, M, OC, IC, stream)
                     : launch_gemv<__half, false>(x, qweight, scales, zeros, bias, out, M, OC, IC, stream);
}

extern "C" __attribute__((visibility("default"))) int nb200_gemv_awq(int dtype, const void *x, const void *qweight, const void *scales, const void *zeros,
                                                                     void *out, int M, int OC, int IC, int group_size, void *stream_) {
    return nb200_gemv_awq_fused(dtype, x, qweight, scales, zeros, nullptr, out, M, OC, IC, group_size, 0, stream_);
}
