// Depthwise 3x3 convolution, NHWC, stride 1, zero padding 1 -- SANA's GLUMBConv.
//
// Replaces dwconv_f16 (src/kernels/dwconv.cu:202-340, a CUTLASS direct-convolution instantiation; module DWCONV, src/Linear.cpp:541-551):
// input [N, H, W, C], weight [C, 3, 3, 1], optional bias [C], output [N, H, W, C], all hT.  HBM bound (9 MACs per element): a thread owns 8
// consecutive channels (one 16-byte vector: adjacent threads read adjacent vectors of the NHWC row) and walks along W with the 3 x 3 window
// in registers, so each output costs three new 16-byte loads that are neighbours' loads one and two steps earlier (L1 / L2 hits) -- every
// input element leaves HBM once, every output element is written once.  The 72 weights of the thread's channels sit in registers as hT pairs.
// Arithmetic: the reference accumulates in the 16-bit type (ElementAccumulator = half_t, dwconv.cu:222-226); here products and the sum over
// the 9 taps (row-major tap order) + bias are fp32 with ONE rounding to hT -- strictly closer to exact math, parity to hT noise
// (tests/test_gpu_dwconv.py against an fp64 convolution).
#include "common.cuh"
#include "ptx.cuh"

namespace nb200 {
namespace {

constexpr int kDwThreads = 128;

template <typename hT>
__global__ void __launch_bounds__(kDwThreads) dwconv3x3_kernel(const hT *__restrict__ x, const hT *__restrict__ w, const hT *__restrict__ bias,
                                                                hT *__restrict__ out, int N, int H, int W, int C) {
    using Tr = HalfTraits<hT>;
    using T2 = typename Tr::T2;
    ptx::griddep_launch_dependents();
    ptx::griddep_wait();
    const int cvecs = C >> 3;
    const long long unit = static_cast<long long>(blockIdx.x) * kDwThreads + threadIdx.x;   // (n, h, channel vector)
    if (unit >= static_cast<long long>(N) * H * cvecs) return;
    const int cv = static_cast<int>(unit % cvecs);
    const int h = static_cast<int>((unit / cvecs) % H);
    const int n = static_cast<int>(unit / (static_cast<long long>(cvecs) * H));
    const int c0 = cv * 8;
    // weights of channels c0 .. c0+7: w[c][r][s]
    float wf[9][8];
#pragma unroll
    for (int k = 0; k < 8; k++)
#pragma unroll
        for (int t = 0; t < 9; t++) wf[t][k] = Tr::to_float(w[static_cast<size_t>(c0 + k) * 9 + t]);
    float bf[8];
#pragma unroll
    for (int k = 0; k < 8; k++) bf[k] = bias != nullptr ? Tr::to_float(bias[c0 + k]) : 0.f;
    const size_t row_pitch = static_cast<size_t>(W) * C;
    const hT *img = x + static_cast<size_t>(n) * H * row_pitch + c0;
    auto load = [&](const int hh, const int ww) -> uint4 {
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) return make_uint4(0, 0, 0, 0);
        return *reinterpret_cast<const uint4 *>(img + static_cast<size_t>(hh) * row_pitch + static_cast<size_t>(ww) * C);
    };
    uint4 win[3][3];   // [row h-1..h+1][column w-1..w+1]
#pragma unroll
    for (int r = 0; r < 3; r++) {
        win[r][1] = load(h + r - 1, -1);
        win[r][2] = load(h + r - 1, 0);
    }
    hT *orow = out + (static_cast<size_t>(n) * H + h) * row_pitch + c0;
    for (int ww = 0; ww < W; ww++) {
#pragma unroll
        for (int r = 0; r < 3; r++) {
            win[r][0] = win[r][1];
            win[r][1] = win[r][2];
            win[r][2] = load(h + r - 1, ww + 1);
        }
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0.f;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int s2 = 0; s2 < 3; s2++) {
                const uint32_t xw[4] = {win[r][s2].x, win[r][s2].y, win[r][s2].z, win[r][s2].w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float2 f = Tr::to_float2(*reinterpret_cast<const T2 *>(&xw[e]));
                    acc[2 * e] = fmaf(f.x, wf[r * 3 + s2][2 * e], acc[2 * e]);
                    acc[2 * e + 1] = fmaf(f.y, wf[r * 3 + s2][2 * e + 1], acc[2 * e + 1]);
                }
            }
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const T2 hv = Tr::from_float2(make_float2(acc[2 * e] + bf[2 * e], acc[2 * e + 1] + bf[2 * e + 1]));
            o[e] = *reinterpret_cast<const uint32_t *>(&hv);
        }
        *reinterpret_cast<uint4 *>(orow + static_cast<size_t>(ww) * C) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace
}  // namespace nb200

// x / out hT [N, H, W, C] contiguous (NHWC), weight hT [C, 3, 3] (the reference's [C, 3, 3, 1]), bias hT [C] or NULL.  C % 8 == 0.
extern "C" __attribute__((visibility("default"))) int nb200_dwconv3x3(int dtype, const void *x, const void *weight, const void *bias, void *out, int N,
                                                                      int H, int W, int C, void *stream_) {
    using namespace nb200;
    reset_launch_count();
    NB200_REQUIRE(x && weight && out, "NULL tensor");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "dtype must be fp16 or bf16");
    NB200_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "C must be a positive multiple of 8");
    NB200_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "x / out must be 16-byte aligned");
    NB200_REQUIRE(x != out, "in-place convolution is not supported");
    if (N == 0) return NB200_OK;
    if (int rc = nb200_check_device()) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long units = static_cast<long long>(N) * H * (C / 8);
    LaunchCfg lc(dim3(static_cast<unsigned>((units + kDwThreads - 1) / kDwThreads)), dim3(kDwThreads), 0, stream);
    if (dtype == NB200_BF16) {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, dwconv3x3_kernel<__nv_bfloat16>, static_cast<const __nv_bfloat16 *>(x), static_cast<const __nv_bfloat16 *>(weight),
                                            static_cast<const __nv_bfloat16 *>(bias), static_cast<__nv_bfloat16 *>(out), N, H, W, C));
    } else {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, dwconv3x3_kernel<__half>, static_cast<const __half *>(x), static_cast<const __half *>(weight),
                                            static_cast<const __half *>(bias), static_cast<__half *>(out), N, H, W, C));
    }
    count_launch();
    return NB200_OK;
}
