// Per-head RMSNorm(Q, K) + rotary embedding applied IN PLACE to a [M, 3*H*128] hT projection -- the second half of the QKV linear when
// the GEMM runs its plain epilogue (gemm_nvfp4_cluster.cu: dispatch, "split" route).
//
// Why it exists (r02 launch list, profiles/r02f_launches.md): with the RoPE epilogue fused, every epilogue thread (= one output row) needs
// its row's 64 (sin, cos) pairs = 512 B of the rotary table for every 128-wide head it finishes -- 128 KB of extra ingest per 256 x 256 tile
// on a kernel whose floor is the SM's inbound port, fetched as 64 uncoalesced 8-byte loads per thread.  The fused launch took 115.6 us for
// 4352 x 3072 -> 9216 against 60.6 us for the same GEMM with the plain epilogue.  Here a WARP owns one row and its lanes own heads, so the
// table is read once per row as warp-wide broadcasts and the arithmetic is the fused epilogue's, instruction for instruction:
//   y (already rounded to hT by the GEMM's store, as the reference rounds fpsum: epilogues.cuh:327-341)
//   sumsq over the head's 128 columns in column order (fp32 FMA chain), coef = rsqrt.approx.ftz(sumsq / 128 + 1e-6)
//   x = y * (coef * w[c]);  (y0, y1) <- (x0 cos - x1 sin, x0 sin + x1 cos)   [epilogues.cuh:343-367], one rounding to hT
// so both routes give bit-identical outputs (tests/test_gpu_fused.py).  V columns are not touched.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace nb200 {
namespace {

constexpr int kRopeWarps = 8;

template <typename hT>
__global__ void __launch_bounds__(kRopeWarps * 32) rope_inplace_kernel(hT *__restrict__ qkv, int M, int N, const hT *__restrict__ norm_q,
                                                                        const hT *__restrict__ norm_k, const float *__restrict__ rotary) {
    using Tr = HalfTraits<hT>;
    using T2 = typename Tr::T2;
    __shared__ float normw[256];   // q | k
    ptx::griddep_launch_dependents();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) normw[i] = Tr::to_float((i < 128 ? norm_q : norm_k)[i & 127]);
    __syncthreads();
    ptx::griddep_wait();   // qkv is the GEMM's output
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int heads_qk = 2 * (N / 3) / 128;            // Q heads then K heads, contiguous in the row
    const int groups = (heads_qk + 31) / 32;           // lanes = heads, 32 at a time
    const long long unit = static_cast<long long>(blockIdx.x) * kRopeWarps + warp;
    if (unit >= static_cast<long long>(M) * groups) return;
    const int m = static_cast<int>(unit / groups), head = static_cast<int>(unit % groups) * 32 + lane;
    if (head >= heads_qk) return;
    const int part = head >= heads_qk / 2;              // 0 = Q, 1 = K
    const float *w = normw + part * 128;
    hT *row = qkv + static_cast<size_t>(m) * N + static_cast<size_t>(head) * 128;
    // reference pack_rotemb order (transformer_flux.py:60-92): float index of (row m, pair pr, sin|cos)
    //   ((((m/16*16 + pr/4)*8 + m%8)*4 + pr%4)*2 + (m%16)/8)*2 + {0,1}   -- the same for every lane of the warp: broadcast loads
    const float *rot_row = rotary + (static_cast<size_t>(m >> 4) * 16 * 8 + (m & 7)) * 16 + ((m >> 3) & 1) * 2;

    uint4 v[16];
#pragma unroll
    for (int c = 0; c < 16; c++) v[c] = *reinterpret_cast<const uint4 *>(row + c * 8);
    float sumsq = 0.f;
#pragma unroll
    for (int c = 0; c < 16; c++) {
        const uint32_t xw[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float2 r = Tr::to_float2(*reinterpret_cast<const T2 *>(&xw[e]));
            sumsq = fmaf(r.x, r.x, sumsq);
            sumsq = fmaf(r.y, r.y, sumsq);
        }
    }
    const float coef = rsqrt_approx_ftz(sumsq / 128.f + 1e-6f);
#pragma unroll
    for (int c = 0; c < 16; c++) {
        const uint32_t xw[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int hc = c * 8 + 2 * e, pr = hc >> 1;
            const float2 r = Tr::to_float2(*reinterpret_cast<const T2 *>(&xw[e]));
            const float x0 = r.x * (coef * w[hc]);
            const float x1 = r.y * (coef * w[hc + 1]);
            const float2 sc = __ldg(reinterpret_cast<const float2 *>(rot_row + (pr >> 2) * 128 + (pr & 3) * 4));   // (sin, cos)
            float y0 = x0 * sc.y - x1 * sc.x;
            float y1 = x0 * sc.x + x1 * sc.y;
            if constexpr (!Tr::kIsBf16) {   // fp16 stores clamp (gemm_base.cuh:688-696)
                y0 = fminf(fmaxf(y0, -65504.f), 65504.f);
                y1 = fminf(fmaxf(y1, -65504.f), 65504.f);
            }
            const T2 hv = Tr::from_float2(make_float2(y0, y1));
            o[e] = *reinterpret_cast<const uint32_t *>(&hv);
        }
        *reinterpret_cast<uint4 *>(row + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}


// The same arithmetic with the PackQKV hand-off (EpiloguePackQKV, epilogues.cuh:427-550): reads the plain projection [Mp, 3*H*128] hT and writes
// the attention kernel's operands -- fp16 [heads][rows][128] each, Q and K normalised + rotated, V converted, rows >= attn_tokens masked
// 0 / NaN / 0 (epilogues.cuh:479-489,539-545) -- exactly what the fused epilogue of the GEMM writes (hT rounding of the rotated value,
// then hT -> fp16 through fp32).  Lanes = heads of all three parts.
template <typename hT>
__global__ void __launch_bounds__(kRopeWarps * 32) rope_pack_kernel(const hT *__restrict__ qkv, int Mp, int N, const hT *__restrict__ norm_q,
                                                                     const hT *__restrict__ norm_k, const float *__restrict__ rotary, __half *out_q,
                                                                     __half *out_k, __half *out_v, long long sq, long long sk, long long sv, int attn_tokens) {
    using Tr = HalfTraits<hT>;
    using T2 = typename Tr::T2;
    __shared__ float normw[256];   // q | k
    ptx::griddep_launch_dependents();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) normw[i] = Tr::to_float((i < 128 ? norm_q : norm_k)[i & 127]);
    __syncthreads();
    ptx::griddep_wait();   // qkv is the GEMM's output
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = (N / 3) / 128;
    const int groups = (3 * H + 31) / 32;
    const long long unit = static_cast<long long>(blockIdx.x) * kRopeWarps + warp;
    if (unit >= static_cast<long long>(Mp) * groups) return;
    const int m = static_cast<int>(unit / groups), head = static_cast<int>(unit % groups) * 32 + lane;
    if (head >= 3 * H) return;
    const int part = head / H, hh = head % H;           // 0 = Q, 1 = K, 2 = V
    const hT *row = qkv + static_cast<size_t>(m) * N + static_cast<size_t>(head) * 128;
    __half *dst = (part == 0 ? out_q + hh * sq : part == 1 ? out_k + hh * sk : out_v + hh * sv) + static_cast<size_t>(m) * 128;
    if (m >= attn_tokens) {   // pad rows: the key mask is NaN, queries and values are zero
        const uint32_t fill = part == 1 ? 0x7FFF7FFFu : 0u;
#pragma unroll
        for (int c = 0; c < 16; c++) *reinterpret_cast<uint4 *>(dst + c * 8) = make_uint4(fill, fill, fill, fill);
        return;
    }
    uint4 v[16];
#pragma unroll
    for (int c = 0; c < 16; c++) v[c] = *reinterpret_cast<const uint4 *>(row + c * 8);
    float coef = 0.f;
    const float *w = normw + (part & 1) * 128;
    const float *rot_row = rotary + (static_cast<size_t>(m >> 4) * 16 * 8 + (m & 7)) * 16 + ((m >> 3) & 1) * 2;
    if (part < 2) {
        float sumsq = 0.f;
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const uint32_t xw[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 r = Tr::to_float2(*reinterpret_cast<const T2 *>(&xw[e]));
                sumsq = fmaf(r.x, r.x, sumsq);
                sumsq = fmaf(r.y, r.y, sumsq);
            }
        }
        coef = rsqrt_approx_ftz(sumsq / 128.f + 1e-6f);
    }
#pragma unroll
    for (int c = 0; c < 16; c++) {
        const uint32_t xw[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            T2 hv = *reinterpret_cast<const T2 *>(&xw[e]);
            if (part < 2) {
                const int hc = c * 8 + 2 * e, pr = hc >> 1;
                const float2 r = Tr::to_float2(hv);
                const float x0 = r.x * (coef * w[hc]);
                const float x1 = r.y * (coef * w[hc + 1]);
                const float2 sc = __ldg(reinterpret_cast<const float2 *>(rot_row + (pr >> 2) * 128 + (pr & 3) * 4));   // (sin, cos)
                float y0 = x0 * sc.y - x1 * sc.x;
                float y1 = x0 * sc.x + x1 * sc.y;
                if constexpr (!Tr::kIsBf16) {
                    y0 = fminf(fmaxf(y0, -65504.f), 65504.f);
                    y1 = fminf(fmaxf(y1, -65504.f), 65504.f);
                }
                hv = Tr::from_float2(make_float2(y0, y1));
            }
            const __half2 hh2 = __float22half2_rn(Tr::to_float2(hv));   // hT -> fp16 through fp32 (epilogues.cuh:446-453)
            o[e] = *reinterpret_cast<const uint32_t *>(&hh2);
        }
        *reinterpret_cast<uint4 *>(dst + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace

// qkv hT [M, N] (N = 3 * heads * 128), rotary: the reference's packed table covering at least the rows < M
int rope_inplace_dispatch(int dtype, void *qkv, int M, int N, const void *norm_q, const void *norm_k, const float *rotary, cudaStream_t stream) {
    NB200_REQUIRE(qkv && norm_q && norm_k && rotary, "NULL tensor");
    NB200_REQUIRE(N % 384 == 0, "N must be 3 * heads * 128");
    if (M <= 0) return NB200_OK;
    const int heads_qk = 2 * (N / 3) / 128;
    const long long units = static_cast<long long>(M) * ((heads_qk + 31) / 32);
    const unsigned grid = static_cast<unsigned>((units + kRopeWarps - 1) / kRopeWarps);
    LaunchCfg lc(dim3(grid), dim3(kRopeWarps * 32), 0, stream);
    if (dtype == NB200_BF16) {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rope_inplace_kernel<__nv_bfloat16>, static_cast<__nv_bfloat16 *>(qkv), M, N,
                                            static_cast<const __nv_bfloat16 *>(norm_q), static_cast<const __nv_bfloat16 *>(norm_k), rotary));
    } else {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rope_inplace_kernel<__half>, static_cast<__half *>(qkv), M, N, static_cast<const __half *>(norm_q),
                                            static_cast<const __half *>(norm_k), rotary));
    }
    count_launch();
    return NB200_OK;
}

}  // namespace nb200

namespace nb200 {

// qkv hT [Mp, N] (the plain projection, all Mp rows) -> out_q / out_k / out_v fp16 [heads][>= Mp rows][128] with head pitches sq / sk / sv elements
int rope_pack_dispatch(int dtype, const void *qkv, int Mp, int N, const void *norm_q, const void *norm_k, const float *rotary, void *out_q, void *out_k,
                       void *out_v, long long sq, long long sk, long long sv, int attn_tokens, cudaStream_t stream) {
    NB200_REQUIRE(qkv && norm_q && norm_k && rotary && out_q && out_k && out_v, "NULL tensor");
    NB200_REQUIRE(N % 384 == 0, "N must be 3 * heads * 128");
    if (Mp <= 0) return NB200_OK;
    const int H = (N / 3) / 128;
    const long long units = static_cast<long long>(Mp) * ((3 * H + 31) / 32);
    const unsigned grid = static_cast<unsigned>((units + kRopeWarps - 1) / kRopeWarps);
    LaunchCfg lc(dim3(grid), dim3(kRopeWarps * 32), 0, stream);
    if (dtype == NB200_BF16) {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rope_pack_kernel<__nv_bfloat16>, static_cast<const __nv_bfloat16 *>(qkv), Mp, N,
                                            static_cast<const __nv_bfloat16 *>(norm_q), static_cast<const __nv_bfloat16 *>(norm_k), rotary,
                                            static_cast<__half *>(out_q), static_cast<__half *>(out_k), static_cast<__half *>(out_v), sq, sk, sv, attn_tokens));
    } else {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rope_pack_kernel<__half>, static_cast<const __half *>(qkv), Mp, N, static_cast<const __half *>(norm_q),
                                            static_cast<const __half *>(norm_k), rotary, static_cast<__half *>(out_q), static_cast<__half *>(out_k),
                                            static_cast<__half *>(out_v), sq, sk, sv, attn_tokens));
    }
    count_launch();
    return NB200_OK;
}

}  // namespace nb200
