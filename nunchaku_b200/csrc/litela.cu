// SANA linear attention pieces of the QKV projection:
//   EpilogueLiteLA            src/kernels/zgemm/epilogues.cuh:552-691 (wiring gemm_w4a4_launch_impl.cuh:311-346)
//   vk_mul_q_kernel           src/kernels/zgemm/epilogues.cuh:693-760 (launcher gemm_w4a4_launch_impl.cuh:427-448)
//
// The reference computes relu(Q) and the per-head [33 x 32] state  vk[v][k] = sum_t V[t,v] * relu(K[t,k])
// (row 32: sum_t relu(K[t,k])) inside the GEMM epilogue, and so does gemm_w4a4.cu's EPI_LITELA (NVFP4, N / 3 a multiple of 128: DESIGN.md
// section 4.3).  This file is the other route -- INT4 (measured faster split) and odd shapes: the GEMM writes the plain hT tile and
// `litela_vk_kernel` consumes it while it is still L2-resident (SANA: [B*1024, 6912] bf16 = 28 MB against a
// 126 MB L2); same arithmetic: relu and the operands in hT,
// products and sums in fp32, partial sums of token blocks combined with fp32 atomics (reference: reduce_add
// per 256-token block).  Channel layout of the projection output (N = 3 * heads * 32): [ Q (N/3) | per head:
// K (32), V (32) ].
#include "common.cuh"

namespace nb200 {
namespace {

constexpr int kHeadDim = 32;          // LITELA_HEAD_DIM
constexpr int kVkThreads = 288;       // 9 warps: thread -> (output row 0..32, group of 4 columns)
constexpr int kTokTile = 64;          // tokens staged in shared memory per step

// grid (heads, batch, token splits)
template <typename hT>
__global__ void __launch_bounds__(kVkThreads) litela_vk_kernel(const hT *__restrict__ qkv, float *__restrict__ out_vk, int tokens,
                                                                int tokens_per_split, int N, int heads) {
    using Tr = HalfTraits<hT>;
    __shared__ float sk[kTokTile][kHeadDim];      // relu(K) as fp32
    __shared__ float sv[kTokTile][kHeadDim + 1];  // V as fp32, column 32 = 1 (the "sum of K" row)
    const int head = blockIdx.x, b = blockIdx.y;
    const int t_begin = blockIdx.z * tokens_per_split;
    const int t_end = min(tokens, t_begin + tokens_per_split);
    const int vrow = threadIdx.x >> 3;            // 0..35 (rows >= 33 idle)
    const int k0 = (threadIdx.x & 7) * 4;
    const hT *base = qkv + static_cast<size_t>(b) * tokens * N + N / 3 + head * 2 * kHeadDim;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t0 = t_begin; t0 < t_end; t0 += kTokTile) {
        __syncthreads();
        // stage 64 tokens x (32 K + 32 V): 8 hT (16 bytes) per thread-iteration
        for (int i = threadIdx.x; i < kTokTile * 8; i += kVkThreads) {
            const int tt = i >> 3, c8 = i & 7;    // c8 0..3 -> K, 4..7 -> V
            const int t = t0 + tt;
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (t < t_end) raw = *reinterpret_cast<const uint4 *>(base + static_cast<size_t>(t) * N + c8 * 8);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = Tr::to_float2(*reinterpret_cast<const typename Tr::T2 *>(&w[e]));
                if (c8 < 4) {  // relu on the hT value (epilogues.cuh:611-613)
                    sk[tt][c8 * 8 + 2 * e] = fmaxf(f.x, 0.f);
                    sk[tt][c8 * 8 + 2 * e + 1] = fmaxf(f.y, 0.f);
                } else {
                    sv[tt][(c8 - 4) * 8 + 2 * e] = f.x;
                    sv[tt][(c8 - 4) * 8 + 2 * e + 1] = f.y;
                }
            }
            if (c8 == 0) sv[tt][kHeadDim] = t < t_end ? 1.f : 0.f;
        }
        __syncthreads();
        if (vrow <= kHeadDim) {
#pragma unroll 8
            for (int tt = 0; tt < kTokTile; tt++) {
                const float v = sv[tt][vrow];
                const float4 k = *reinterpret_cast<const float4 *>(&sk[tt][k0]);
                acc[0] = fmaf(v, k.x, acc[0]);
                acc[1] = fmaf(v, k.y, acc[1]);
                acc[2] = fmaf(v, k.z, acc[2]);
                acc[3] = fmaf(v, k.w, acc[3]);
            }
        }
    }
    if (vrow <= kHeadDim) {
        float *dst = out_vk + ((static_cast<size_t>(b) * heads + head) * (kHeadDim + 1) + vrow) * kHeadDim + k0;
#pragma unroll
        for (int i = 0; i < 4; i++) atomicAdd(dst + i, acc[i]);
    }
}

// out_q[b, t, c] = relu(qkv[b, t, c]) for c < N/3   (epilogues.cuh:676-688)
template <typename hT>
__global__ void __launch_bounds__(256) litela_relu_q_kernel(const hT *__restrict__ qkv, hT *__restrict__ out_q, long long rows, int N) {
    using Tr = HalfTraits<hT>;
    const int nq8 = N / 3 / 8;
    const long long total = rows * nq8;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
        const long long r = i / nq8;
        const int c8 = static_cast<int>(i % nq8);
        uint4 raw = *reinterpret_cast<const uint4 *>(qkv + r * N + c8 * 8);
        uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            typename Tr::T2 h = *reinterpret_cast<typename Tr::T2 *>(&w[e]);
            typename Tr::T2 z;
            z.x = Tr::from_float(0.f);
            z.y = z.x;
            h = __hmax2(h, z);
            w[e] = *reinterpret_cast<uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(out_q + r * (N / 3) + c8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// q[b, t, head, :] <- (q . vk[0..31]) / (q . vk[32] + eps), one thread per (token, head)  (epilogues.cuh:693-760)
template <typename hT>
__global__ void __launch_bounds__(128) vk_mul_q_kernel(hT *__restrict__ q, const float *__restrict__ vk, float eps, int tokens, int heads) {
    using Tr = HalfTraits<hT>;
    __shared__ float svk[(kHeadDim + 1) * kHeadDim];
    const int head = blockIdx.y, b = blockIdx.z;
    const float *lvk = vk + (static_cast<size_t>(b) * heads + head) * (kHeadDim + 1) * kHeadDim;
    for (int i = threadIdx.x; i < (kHeadDim + 1) * kHeadDim; i += blockDim.x) svk[i] = lvk[i];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tokens) return;
    hT *lq = q + ((static_cast<size_t>(b) * tokens + t) * heads + head) * kHeadDim;
    float qf[kHeadDim];
#pragma unroll
    for (int i = 0; i < kHeadDim; i += 8) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(lq + i);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float2 f = Tr::to_float2(*reinterpret_cast<const typename Tr::T2 *>(&w[e]));
            qf[i + 2 * e] = f.x;
            qf[i + 2 * e + 1] = f.y;
        }
    }
    float den = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadDim; i++) den = fmaf(qf[i], svk[kHeadDim * kHeadDim + i], den);
    den += eps;
#pragma unroll 1
    for (int j0 = 0; j0 < kHeadDim; j0 += 8) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float o[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int j = j0 + 2 * e + u;
                float a = 0.f;
#pragma unroll
                for (int i = 0; i < kHeadDim; i++) a = fmaf(qf[i], svk[j * kHeadDim + i], a);   // same i order as the reference loop
                o[u] = __fdividef(a, den);
            }
            const typename Tr::T2 h = Tr::from_float2(make_float2(o[0], o[1]));
            w[e] = *reinterpret_cast<const uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(lq + j0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace
}  // namespace nb200

using namespace nb200;

#pragma GCC visibility push(default)
extern "C" {

int nb200_litela_vk(int dtype, const void *qkv, void *out_q, float *out_vk, int batch, int tokens, int N, void *stream_) {
    NB200_REQUIRE(qkv != nullptr && out_q != nullptr && out_vk != nullptr, "litela_vk: null tensor");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "litela_vk: dtype must be fp16 or bf16");
    NB200_REQUIRE(batch > 0 && tokens > 0 && N > 0 && N % (3 * kHeadDim) == 0, "litela_vk: N must be 3 * heads * 32");
    NB200_REQUIRE(N % 24 == 0, "litela_vk: N / 3 must be a multiple of 8");
    NB200_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out_q)) & 15) == 0, "litela_vk: 16-byte alignment");
    auto stream = static_cast<cudaStream_t>(stream_);
    const int heads = N / (3 * kHeadDim);
    // the reference zero-fills out_vk inside the op (launch_impl:336)
    NB200_CUDA_CHECK(cudaMemsetAsync(out_vk, 0, static_cast<size_t>(batch) * heads * (kHeadDim + 1) * kHeadDim * sizeof(float), stream));
    int splits = (tokens + 511) / 512;
    const int per = ((tokens + splits - 1) / splits + kTokTile - 1) / kTokTile * kTokTile;
    splits = (tokens + per - 1) / per;
    const long long rows = static_cast<long long>(batch) * tokens;
    long long g = (rows * (N / 3 / 8) + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (dtype == NB200_BF16) {
        litela_relu_q_kernel<__nv_bfloat16><<<static_cast<int>(g), 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(qkv),
                                                                                    static_cast<__nv_bfloat16 *>(out_q), rows, N);
        litela_vk_kernel<__nv_bfloat16><<<dim3(heads, batch, splits), kVkThreads, 0, stream>>>(static_cast<const __nv_bfloat16 *>(qkv), out_vk,
                                                                                             tokens, per, N, heads);
    } else {
        litela_relu_q_kernel<__half><<<static_cast<int>(g), 256, 0, stream>>>(static_cast<const __half *>(qkv), static_cast<__half *>(out_q), rows, N);
        litela_vk_kernel<__half><<<dim3(heads, batch, splits), kVkThreads, 0, stream>>>(static_cast<const __half *>(qkv), out_vk, tokens, per, N,
                                                                                      heads);
    }
    NB200_CUDA_CHECK(cudaGetLastError());
    count_launch(2);
    return NB200_OK;
}

int nb200_linearattn_vk_mul_q(int dtype, void *q, const float *vk, int batch, int tokens, int heads, float eps, void *stream_) {
    NB200_REQUIRE(q != nullptr && vk != nullptr, "linearattn_vk_mul_q: null tensor");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "linearattn_vk_mul_q: dtype must be fp16 or bf16");
    NB200_REQUIRE(batch > 0 && tokens > 0 && heads > 0, "linearattn_vk_mul_q: bad sizes");
    NB200_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "linearattn_vk_mul_q: q must be 16-byte aligned");
    auto stream = static_cast<cudaStream_t>(stream_);
    const dim3 grid((tokens + 127) / 128, heads, batch);
    if (dtype == NB200_BF16)
        vk_mul_q_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(static_cast<__nv_bfloat16 *>(q), vk, eps, tokens, heads);
    else
        vk_mul_q_kernel<__half><<<grid, 128, 0, stream>>>(static_cast<__half *>(q), vk, eps, tokens, heads);
    NB200_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return NB200_OK;
}

}  // extern "C"
#pragma GCC visibility pop
