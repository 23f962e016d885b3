// One-time (parameter-load) repack of the reference's mma.sync-fragment-ordered tensors into the
// B200 layouts of include/nunchaku_b200.h.  Pure gathers driven by the closed-form index maps of
// the reference formats (nunchaku/lora/flux/packer.py:187-437).
// No reference twin exists: the reference consumes its own layout directly
// (GEMM_W4A4::loadParam, src/Linear.cpp:124-154).
#include "common.cuh"

namespace nb200 {
namespace {

constexpr int kThreads = 256;

__host__ __device__ inline int ceil_div_i(long long a, int b) { return static_cast<int>((a + b - 1) / b); }

// reference flat index of scale (n, grp) among G groups: [nt][grp][lane = a*4 + c2][b*2 + d],
// n = nt*128 + a*16 + b*8 + c2*2 + d            (packer.py:241-301)
__device__ __forceinline__ size_t ref_scale_index(int n, int grp, int G) {
    const int nt = n >> 7, ni = n & 127;
    const int a = ni >> 4, b = (ni >> 3) & 1, c2 = (ni >> 1) & 3, d = ni & 1;
    return ((static_cast<size_t>(nt) * G + grp) * 32 + (a * 4 + c2)) * 4 + (b * 2 + d);
}

// reference flat index of low-rank element; `c16` indexes the 16-blocks of the stored first
// dimension (N for up, K for down), see packer.py:362-398:
//   [C/16][R/16][lane = g*4 + t][h][c][e]
//   up  : n = 16*i + h*8 + g, r = 16*u + c*8 + t*2 + e
//   down: r = 16*u + h*8 + g, k = 16*i + c*8 + t*2 + e
__device__ __forceinline__ size_t ref_lowrank_index(int i, int u, int g, int t, int h, int c, int e, int R) {
    return ((((static_cast<size_t>(i) * (R >> 4) + u) * 32 + (g * 4 + t)) * 2 + h) * 2 + c) * 2 + e;
}

// ---- qweight --------------------------------------------------------------------------------
// The 8 nibbles of k = 8q .. 8q+7 of one row share a source word (A.1): only the word address
// and (INT4) the nibble order / offset change.
__global__ void repack_qweight_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int N, int K,
                                      int fp4) {
    const size_t total = static_cast<size_t>(N) * (K >> 3);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int n = static_cast<int>(i / (K >> 3));
        const int k = static_cast<int>(i % (K >> 3)) << 3;
        const int nt = n >> 7, ni = n & 127;
        const int j = ni >> 4, h = (ni >> 3) & 1, g = ni & 7;
        const int kt = k >> 6, ki = k & 63;
        const int c = ki >> 5, t = (ki >> 3) & 3;
        const size_t word = (((static_cast<size_t>(nt) * (K >> 6) + kt) * 8 + j) * 32 + (g * 4 + t)) * 4 + (h * 2 + c);
        uint32_t w = src[word];
        if (!fp4) {
            // nibble p <- e(2p), nibble p+4 <- e(2p+1); two's complement -> offset binary
            uint32_t o = 0;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                o |= ((w >> (8 * p)) & 0xFu) << (4 * p);
                o |= ((w >> (8 * p + 4)) & 0xFu) << (4 * (p + 4));
            }
            w = o ^ 0x88888888u;
        }
        dst[i] = w;
    }
}

// ---- scales -----------------------------------------------------------------------------------
__global__ void repack_wscales_int4_kernel(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int N, int G) {
    const size_t total = static_cast<size_t>(N) * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int grp = static_cast<int>(i / N), n = static_cast<int>(i % N);
        dst[i] = src[ref_scale_index(n, grp, G)];
    }
}

// reference micro-scale flat index: [nt][kt][lane = s*4 + q][p][kk], n = nt*128 + p*32 + q*8 + s,
// g16 = kt*4 + kk  (packer.py:303-360)
__global__ void repack_wscales_fp4_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int N, int K) {
    const int G16 = K >> 4, KT = K >> 6;
    const size_t total = static_cast<size_t>(N) * G16;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        // iterate in destination order
        const size_t tile = i >> 9;        // 512-byte tile = (n128, kt)
        const int within = static_cast<int>(i & 511);
        const int n128 = static_cast<int>(tile / KT), kt = static_cast<int>(tile % KT);
        const int r32 = within >> 4, q4 = (within >> 2) & 3, kk = within & 3;
        const int n = n128 * 128 + q4 * 32 + r32;
        const int ni = n & 127;
        const int p = ni >> 5, q = (ni >> 3) & 3, s = ni & 7;
        const size_t sidx = ((((static_cast<size_t>(n128) * KT + kt) * 32 + (s * 4 + q)) * 4 + p) * 4) + kk;
        dst[i] = src[sidx];
    }
}

template <typename hT>
__global__ void repack_channel_vector_kernel(const hT *__restrict__ src, void *__restrict__ dst, int N, int out_f32,
                                             float mul) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const hT v = src[ref_scale_index(n, 0, 1)];
        if (out_f32)
            reinterpret_cast<float *>(dst)[n] = HalfTraits<hT>::to_float(v) * mul;
        else
            reinterpret_cast<hT *>(dst)[n] = v;
    }
}

// ---- low-rank factors ---------------------------------------------------------------------------
// dst blocks [Rp/32][N/8][4][8 rows][8 ranks]: a [rows x 32 ranks] slab of any 8-aligned row range
// is contiguous and already in UMMA no-swizzle K-major core-matrix order
// (LBO = 128 B between the 4 rank-octets, SBO = 512 B between 8-row groups).
template <typename hT>
__global__ void repack_lora_up_kernel(const hT *__restrict__ src, hT *__restrict__ dst, const float *__restrict__ cscale,
                                      int N, int R, int Rp, int *__restrict__ bad_channel) {
    constexpr float kMax = HalfTraits<hT>::kIsBf16 ? 3.38e38f : 65504.f;
    const size_t total = static_cast<size_t>(N) * Rp;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int e8 = static_cast<int>(i & 7), r8 = static_cast<int>((i >> 3) & 7), o = static_cast<int>((i >> 6) & 3);
        const size_t blk = i >> 8;  // chunk * (N/8) + n/8
        const int chunk = static_cast<int>(blk / (N >> 3)), n8 = static_cast<int>(blk % (N >> 3));
        const int n = n8 * 8 + r8;
        const int r = chunk * 32 + o * 8 + e8;
        float v = 0.f;
        if (r < R) {
            const int ii = n >> 4, h = (n >> 3) & 1, g = n & 7;
            const int u = r >> 4, c = (r >> 3) & 1, t = (r >> 1) & 3, e = r & 1;
            v = HalfTraits<hT>::to_float(src[ref_lowrank_index(ii, u, g, t, h, c, e, R)]);
            if (cscale != nullptr) {
                v = v / cscale[n];
                if (!(fabsf(v) <= kMax)) atomicMax(bad_channel, n + 1);   // zero / denormal / non-finite scale, or the quotient overflows hT
            }
        }
        dst[i] = HalfTraits<hT>::from_float(v);
    }
}

// dst [K/32][Rp/8][lane = gq*4 + t][8]: element e of lane (gq, t) = Ld[rank 8j + gq][k = kb*32 + 8t + e]
// (B fragments of the k-permuted mma.sync in quantize.cu)
// second half additionally: the TMA kernel (quantize_v2.cu) feeds ldmatrix fragments in true k order:
// first half dst [K/32][Rp/8][lane][ks2][b][e] = Ld[8j + gq][kb*32 + 16*ks2 + 8*b + 2*t + e]
template <typename hT>
__global__ void repack_lora_down_kernel(const hT *__restrict__ src, hT *__restrict__ dst, int K, int R, int Rp) {
    const size_t total = static_cast<size_t>(K) * Rp;
    for (size_t i2 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i2 < 2 * total;
         i2 += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const bool permuted = i2 >= total;
        const size_t i = permuted ? i2 - total : i2;
        const int e8 = static_cast<int>(i & 7), lane = static_cast<int>((i >> 3) & 31);
        const size_t blk = i >> 8;  // kb * (Rp/8) + j
        const int kb = static_cast<int>(blk / (Rp >> 3)), j = static_cast<int>(blk % (Rp >> 3));
        const int gq = lane >> 2, t = lane & 3;
        const int r = j * 8 + gq;
        const int k = permuted ? kb * 32 + t * 8 + e8 : kb * 32 + (e8 >> 2) * 16 + ((e8 >> 1) & 1) * 8 + t * 2 + (e8 & 1);
        hT v = HalfTraits<hT>::from_float(0.f);
        if (r < R) {
            const int u = r >> 4, h = (r >> 3) & 1, g = r & 7;
            const int ii = k >> 4, c = (k >> 3) & 1, tt = (k >> 1) & 3, e = k & 1;
            v = src[ref_lowrank_index(ii, u, g, tt, h, c, e, R)];
        }
        dst[i2] = v;
    }
}

// dst = logical Ld[r][k] row-major [R][K]: the TMA source of the fused next-layer down projection
template <typename hT>
__global__ void repack_lora_down_rowmajor_kernel(const hT *__restrict__ src, hT *__restrict__ dst, int K, int R) {
    const size_t total = static_cast<size_t>(K) * R;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / K), k = static_cast<int>(i % K);
        const int u = r >> 4, h = (r >> 3) & 1, g = r & 7;
        const int ii = k >> 4, c = (k >> 3) & 1, tt = (k >> 1) & 3, e = k & 1;
        dst[i] = src[ref_lowrank_index(ii, u, g, tt, h, c, e, R)];
    }
}

inline int grid_for(size_t total) {
    size_t b = (total + kThreads - 1) / kThreads;
    return static_cast<int>(b > 148 * 16 ? 148 * 16 : (b == 0 ? 1 : b));
}

}  // namespace
}  // namespace nb200

using namespace nb200;

extern "C" __attribute__((visibility("default"))) int nb200_repack_qweight(const void *src, void *dst, int N, int K, int fp4, void *stream) {
    reset_launch_count();
    NB200_REQUIRE(src && dst, "NULL tensor");
    NB200_REQUIRE(N % 128 == 0 && K % 128 == 0 && N > 0 && K > 0, "N and K must be positive multiples of 128");
    const size_t total = static_cast<size_t>(N) * (K >> 3);
    repack_qweight_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint32_t *>(src), static_cast<uint32_t *>(dst), N, K, fp4);
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    return NB200_OK;
}

extern "C" __attribute__((visibility("default"))) int nb200_repack_wscales_int4(const void *src, void *dst, int N, int K, void *stream) {
    reset_launch_count();
    NB200_REQUIRE(src && dst, "NULL tensor");
    NB200_REQUIRE(N % 128 == 0 && K % 64 == 0, "N % 128 == 0 and K % 64 == 0 required");
    const size_t total = static_cast<size_t>(N) * (K >> 6);
    repack_wscales_int4_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t *>(src), static_cast<uint16_t *>(dst), N, K >> 6);
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    return NB200_OK;
}

extern "C" __attribute__((visibility("default"))) int nb200_repack_wscales_fp4(const void *src, void *dst, int N, int K, void *stream) {
    reset_launch_count();
    NB200_REQUIRE(src && dst, "NULL tensor");
    NB200_REQUIRE(N % 128 == 0 && K % 64 == 0, "N % 128 == 0 and K % 64 == 0 required");
    const size_t total = static_cast<size_t>(N) * (K >> 4);
    repack_wscales_fp4_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t *>(src), static_cast<uint8_t *>(dst), N, K);
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    return NB200_OK;
}

extern "C" __attribute__((visibility("default"))) int nb200_repack_channel_vector(const void *src, void *dst, int N, int dtype, int out_f32, float mul,
                                           void *stream) {
    reset_launch_count();
    NB200_REQUIRE(src && dst, "NULL tensor");
    NB200_REQUIRE(N % 128 == 0 && N > 0, "N must be a positive multiple of 128");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "dtype must be fp16 or bf16");
    if (dtype == NB200_BF16)
        repack_channel_vector_kernel<__nv_bfloat16><<<grid_for(N), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16 *>(src), dst, N, out_f32, mul);
    else
        repack_channel_vector_kernel<__half><<<grid_for(N), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __half *>(src), dst, N, out_f32, mul);
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    return NB200_OK;
}

extern "C" __attribute__((visibility("default"))) int nb200_repack_lora_up(const void *src, void *dst, const float *cscale, int N, int R, int dtype,
                                    void *stream) {
    reset_launch_count();
    NB200_REQUIRE(src && dst, "NULL tensor");
    NB200_REQUIRE(N % 128 == 0 && R % 16 == 0 && R > 0, "N % 128 == 0, R % 16 == 0, R > 0 required");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "dtype must be fp16 or bf16");
    const int Rp = (R + 31) / 32 * 32;
    const size_t total = static_cast<size_t>(N) * Rp;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // The epilogue computes (acc + lora) * cscale + bias on ONE accumulator, so the low-rank factor is stored divided by cscale
    // (DESIGN.md section 4.2).  A channel whose scale is zero, denormal or non-finite, or whose quotient leaves hT's range, cannot be
    // represented that way (the reference adds the low-rank term unscaled): detect it here, at load time, instead of producing inf / NaN
    // at run time.  (Skipped while the stream is being captured: the check needs a device -> host read.)
    int *bad_dev = nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NB200_CUDA_CHECK(cudaStreamIsCapturing(st, &cap));
    const bool validate = cscale != nullptr && cap == cudaStreamCaptureStatusNone;
    if (cscale != nullptr) {
        NB200_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&bad_dev), sizeof(int), st));
        NB200_CUDA_CHECK(cudaMemsetAsync(bad_dev, 0, sizeof(int), st));
    }
    if (dtype == NB200_BF16)
        repack_lora_up_kernel<__nv_bfloat16><<<grid_for(total), kThreads, 0, st>>>(static_cast<const __nv_bfloat16 *>(src), static_cast<__nv_bfloat16 *>(dst),
                                                                                    cscale, N, R, Rp, bad_dev);
    else
        repack_lora_up_kernel<__half><<<grid_for(total), kThreads, 0, st>>>(static_cast<const __half *>(src), static_cast<__half *>(dst), cscale, N, R, Rp,
                                                                             bad_dev);
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    int bad = 0;
    if (validate) {
        NB200_CUDA_CHECK(cudaMemcpyAsync(&bad, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
        NB200_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    if (bad_dev != nullptr) NB200_CUDA_CHECK(cudaFreeAsync(bad_dev, st));
    if (bad != 0)
        return fail(NB200_ERR_UNSUPPORTED, "repack_lora_up: output channel " + std::to_string(bad - 1) +
                                               ": alpha * wcscales is zero / denormal / non-finite, or lora_up / (alpha * wcscales) overflows the 16-bit type; "
                                               "the fused epilogue cannot represent this layer (use bf16, or rescale lora_up / lora_scales)");
    return NB200_OK;
}

extern "C" __attribute__((visibility("default"))) int nb200_repack_lora_down(const void *src, void *dst, int K, int R, int dtype, void *stream) {
    reset_launch_count();
    NB200_REQUIRE(src && dst, "NULL tensor");
    NB200_REQUIRE(K % 32 == 0 && R % 16 == 0 && R > 0, "K % 32 == 0, R % 16 == 0, R > 0 required");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "dtype must be fp16 or bf16");
    const int Rp = (R + 31) / 32 * 32;
    const size_t total = 2 * static_cast<size_t>(K) * Rp;
    if (dtype == NB200_BF16)
        repack_lora_down_kernel<__nv_bfloat16><<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16 *>(src), static_cast<__nv_bfloat16 *>(dst), K, R, Rp);
    else
        repack_lora_down_kernel<__half><<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __half *>(src), static_cast<__half *>(dst), K, R, Rp);
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    return NB200_OK;
}

extern "C" __attribute__((visibility("default"))) int nb200_repack_lora_down_next(const void *src, void *dst, int K, int R, int dtype, void *stream) {
    reset_launch_count();
    NB200_REQUIRE(src && dst, "NULL tensor");
    NB200_REQUIRE(K % 16 == 0 && R % 16 == 0 && R > 0, "K % 16 == 0, R % 16 == 0, R > 0 required");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "dtype must be fp16 or bf16");
    const size_t total = static_cast<size_t>(K) * R;
    if (dtype == NB200_BF16)
        repack_lora_down_rowmajor_kernel<__nv_bfloat16><<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16 *>(src), static_cast<__nv_bfloat16 *>(dst), K, R);
    else
        repack_lora_down_rowmajor_kernel<__half><<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __half *>(src), static_cast<__half *>(dst), K, R);
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    return NB200_OK;
}
