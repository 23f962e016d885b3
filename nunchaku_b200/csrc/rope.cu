// Per-head RMSNorm(Q, K) + rotary embedding on a [M, 3*H*128] hT projection -- the second half of the QKV linear when the GEMM runs its plain
// epilogue (gemm_nvfp4_cluster.cu: dispatch, "split" route): in place, or with the PackQKV hand-off to the attention kernel.
//
// Why it exists (r02 launch list, profiles/r02f_launches.md): with the RoPE epilogue fused, every epilogue thread (= one output row) needs
// its row's 64 (sin, cos) pairs = 512 B of the rotary table for every 128-wide head it finishes -- 128 KB of extra ingest per 256 x 256 tile
// on a kernel whose floor is the SM's inbound port, fetched as 64 uncoalesced 8-byte loads per thread.  The fused launch took 115.6 us for
// 4352 x 3072 -> 9216 against 60.6 us for the same GEMM with the plain epilogue.  The arithmetic is the fused epilogue's, instruction for instruction:
//   y (already rounded to hT by the GEMM's store, as the reference rounds fpsum: epilogues.cuh:327-341)
//   sumsq over the head's 128 columns (8 partial sums of 16 columns, fixed tree), coef = rsqrt.approx.ftz(sumsq / 128 + 1e-6)
//   x = y * (coef * w[c]);  (y0, y1) <- (x0 cos - x1 sin, x0 sin + x1 cos)   [epilogues.cuh:343-367], one rounding to hT
// so all routes give bit-identical outputs (tests/test_gpu_fused.py, tests/test_gpu_fullsize.py).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace nb200 {
namespace {

constexpr int kRopeWarps = 8;
constexpr int kRopeDefaultVariant = 1;

// One kernel for both hand-offs.  A WARP owns one row; 8 lanes share a 128-wide head (16 consecutive columns = 32 bytes each), so one step of
// the warp covers 4 heads = 1 KB of the row, contiguous (r02h launch list: the first version -- lanes = heads, 256-byte stride between lanes --
// took 43.7 us for 4352 x 6144 against an ~18 us traffic floor).  The row's 64 (sin, cos) pairs are read once: a lane keeps the 8 pairs of its
// columns in registers for all heads.  The head's sum of squares is 8 partial sums of 16 columns (sequential fp32 FMA chains) combined by a
// fixed xor tree ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)) -- the order the GEMMs' fused epilogues use as well, so all routes stay bit-identical.
//   PACK = false: RMSNorm(Q, K) + RoPE in place on qkv [M, N]; V untouched.
//   PACK = true : reads the plain projection [Mp, N] and writes the attention operands (EpiloguePackQKV, epilogues.cuh:427-550): fp16
//                 [heads][rows][128] each, Q and K normalised + rotated, V converted (hT -> fp16 through fp32, epilogues.cuh:446-453), rows >=
//                 attn_tokens masked 0 / NaN / 0 (epilogues.cuh:479-489,539-545).
template <typename hT, bool PACK>
__global__ void __launch_bounds__(kRopeWarps * 32) rope_kernel(hT *__restrict__ qkv, int M, int N, const hT *__restrict__ norm_q,
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
    const int m = blockIdx.x * kRopeWarps + warp;
    if (m >= M) return;
    const int H = (N / 3) / 128;
    const int heads = PACK ? 3 * H : 2 * H;            // Q heads, K heads (, V heads): contiguous in the row
    const int sub = lane & 7, hq = lane >> 3;          // 16-column slice of a head, head inside the group of 4
    hT *row = qkv + static_cast<size_t>(m) * N + sub * 16;
    if constexpr (PACK) {
        if (m >= attn_tokens) {   // pad rows: the key mask is NaN, queries and values are zero
            for (int h0 = 0; h0 < heads; h0 += 4) {
                const int head = h0 + hq;
                if (head >= heads) break;
                const int part = head / H, hh = head % H;
                __half *dst = (part == 0 ? out_q + hh * sq : part == 1 ? out_k + hh * sk : out_v + hh * sv) + static_cast<size_t>(m) * 128 + sub * 16;
                const uint32_t fill = part == 1 ? 0x7FFF7FFFu : 0u;
                reinterpret_cast<uint4 *>(dst)[0] = make_uint4(fill, fill, fill, fill);
                reinterpret_cast<uint4 *>(dst)[1] = make_uint4(fill, fill, fill, fill);
            }
            return;
        }
    }
    // reference pack_rotemb order (transformer_flux.py:60-92): float index of (row m, pair pr, sin|cos)
    //   ((((m/16*16 + pr/4)*8 + m%8)*4 + pr%4)*2 + (m%16)/8)*2 + {0,1};  this lane's pairs: pr = 8 sub .. 8 sub + 7
    const float *rot_row = rotary + (static_cast<size_t>(m >> 4) * 16 * 8 + (m & 7)) * 16 + ((m >> 3) & 1) * 2;
    float2 sc[8];   // (sin, cos)
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int pr = sub * 8 + i;
        sc[i] = __ldg(reinterpret_cast<const float2 *>(rot_row + (pr >> 2) * 128 + (pr & 3) * 4));
    }
    constexpr int kBatch = 4;   // head groups in flight: 8 x 16-byte loads per lane
    for (int h0 = 0; h0 < heads; h0 += 4 * kBatch) {
        uint4 v[kBatch][2];
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
            const int head = h0 + 4 * u + hq;
            if (head < heads) {
                v[u][0] = *reinterpret_cast<const uint4 *>(row + static_cast<size_t>(head) * 128);
                v[u][1] = *reinterpret_cast<const uint4 *>(row + static_cast<size_t>(head) * 128 + 8);
            } else {
                v[u][0] = v[u][1] = make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
            const int head = h0 + 4 * u + hq;          // (the 8 lanes of a head agree on `head`: the shuffles below stay inside the group)
            const bool live = head < heads;
            const int part = live ? head / H : 0, hh = live ? head % H : 0;
            const uint32_t xw[8] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w, v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w};
            float sumsq = 0.f;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float2 r = Tr::to_float2(*reinterpret_cast<const T2 *>(&xw[e]));
                sumsq = fmaf(r.x, r.x, sumsq);
                sumsq = fmaf(r.y, r.y, sumsq);
            }
            sumsq += __shfl_xor_sync(0xffffffffu, sumsq, 1);
            sumsq += __shfl_xor_sync(0xffffffffu, sumsq, 2);
            sumsq += __shfl_xor_sync(0xffffffffu, sumsq, 4);
            if (!live) continue;
            const float coef = rsqrt_approx_ftz(sumsq / 128.f + 1e-6f);
            const float *w = normw + (part & 1) * 128 + sub * 16;
            uint32_t o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                T2 hv = *reinterpret_cast<const T2 *>(&xw[e]);
                if (part < 2) {
                    const float2 r = Tr::to_float2(hv);
                    const float x0 = r.x * (coef * w[2 * e]);
                    const float x1 = r.y * (coef * w[2 * e + 1]);
                    float y0 = x0 * sc[e].y - x1 * sc[e].x;
                    float y1 = x0 * sc[e].x + x1 * sc[e].y;
                    if constexpr (!Tr::kIsBf16) {   // fp16 stores clamp (gemm_base.cuh:688-696)
                        y0 = fminf(fmaxf(y0, -65504.f), 65504.f);
                        y1 = fminf(fmaxf(y1, -65504.f), 65504.f);
                    }
                    hv = Tr::from_float2(make_float2(y0, y1));
                }
                if constexpr (PACK) {
                    const __half2 h2 = __float22half2_rn(Tr::to_float2(hv));
                    o[e] = *reinterpret_cast<const uint32_t *>(&h2);
                } else {
                    o[e] = *reinterpret_cast<const uint32_t *>(&hv);
                }
            }
            if constexpr (PACK) {
                __half *dst = (part == 0 ? out_q + hh * sq : part == 1 ? out_k + hh * sk : out_v + hh * sv) + static_cast<size_t>(m) * 128 + sub * 16;
                reinterpret_cast<uint4 *>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<uint4 *>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
            } else {
                hT *dst = row + static_cast<size_t>(head) * 128;
                reinterpret_cast<uint4 *>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<uint4 *>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
    }
}


// Variant 2 (NB200_ROPE=2, an ABLATION: measured slower): 16 lanes per head, 8 consecutive columns = ONE 16-byte access per lane, so a warp
// instruction reads / writes two whole heads = 512 contiguous bytes (full 32-byte sectors in a single request; variant 1's lanes own 32 bytes as
// two 16-byte accesses, i.e. every request touches 32 half-used sectors and every sector is written by two partial stores).  Measured on a B200
// (tools/rope_bench.py, profiles/r02j_rope_bench.json; GEMM + rope back to back at 4352 x 3072 -> 9216): 93.2 us against 89.2 us for variant 1
// (plain GEMM 60.4 us) -- sector efficiency is not what holds the kernel at ~3.7 TB/s; bit-identical outputs.  Same arithmetic and summation order: the
// 16-column partial sum p_j is one sequential FMA chain -- the even lane runs columns 0..7, hands the running sum to its odd neighbour, which
// continues over columns 8..15 -- and the eight p_j are combined by the same xor tree ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)).
template <typename hT, bool PACK>
__global__ void __launch_bounds__(kRopeWarps * 32) rope16_kernel(hT *__restrict__ qkv, int M, int N, const hT *__restrict__ norm_q,
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
    const int m = blockIdx.x * kRopeWarps + warp;
    if (m >= M) return;
    const int H = (N / 3) / 128;
    const int heads = PACK ? 3 * H : 2 * H;
    const int sub = lane & 15, hq = lane >> 4;          // 8-column slice of a head, head inside the pair
    hT *row = qkv + static_cast<size_t>(m) * N + sub * 8;
    if constexpr (PACK) {
        if (m >= attn_tokens) {   // pad rows: the key mask is NaN, queries and values are zero
            for (int h0 = 0; h0 < heads; h0 += 2) {
                const int head = h0 + hq;
                if (head >= heads) break;
                const int part = head / H, hh = head % H;
                __half *dst = (part == 0 ? out_q + hh * sq : part == 1 ? out_k + hh * sk : out_v + hh * sv) + static_cast<size_t>(m) * 128 + sub * 8;
                const uint32_t fill = part == 1 ? 0x7FFF7FFFu : 0u;
                *reinterpret_cast<uint4 *>(dst) = make_uint4(fill, fill, fill, fill);
            }
            return;
        }
    }
    // this lane's pairs: pr = 4 sub .. 4 sub + 3  ->  (pr >> 2) = sub, (pr & 3) = i   (pack_rotemb order as in variant 1)
    const float *rot_row = rotary + (static_cast<size_t>(m >> 4) * 16 * 8 + (m & 7)) * 16 + ((m >> 3) & 1) * 2 + sub * 128;
    float2 sc[4];   // (sin, cos)
#pragma unroll
    for (int i = 0; i < 4; i++) sc[i] = __ldg(reinterpret_cast<const float2 *>(rot_row + i * 4));
    constexpr int kBatch = 8;   // head pairs in flight: 8 x 16-byte loads per lane
    for (int h0 = 0; h0 < heads; h0 += 2 * kBatch) {
        uint4 v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
            const int head = h0 + 2 * u + hq;
            v[u] = head < heads ? *reinterpret_cast<const uint4 *>(row + static_cast<size_t>(head) * 128) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
            const int head = h0 + 2 * u + hq;          // (the 16 lanes of a head agree on `head`: the shuffles below stay inside the group)
            const bool live = head < heads;
            const int part = live ? head / H : 0, hh = live ? head % H : 0;
            const uint32_t xw[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            float2 r[4];
#pragma unroll
            for (int e = 0; e < 4; e++) r[e] = Tr::to_float2(*reinterpret_cast<const T2 *>(&xw[e]));
            float pe = 0.f;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                pe = fmaf(r[e].x, r[e].x, pe);
                pe = fmaf(r[e].y, r[e].y, pe);
            }
            float po = __shfl_up_sync(0xffffffffu, pe, 1);   // odd lanes: the chain of columns 0..7 of their 16-column group
#pragma unroll
            for (int e = 0; e < 4; e++) {
                po = fmaf(r[e].x, r[e].x, po);
                po = fmaf(r[e].y, r[e].y, po);
            }
            po += __shfl_xor_sync(0xffffffffu, po, 2);      // (odd lanes hold p_j, j = sub >> 1; even lanes carry along values nobody reads)
            po += __shfl_xor_sync(0xffffffffu, po, 4);
            po += __shfl_xor_sync(0xffffffffu, po, 8);
            const float sumsq = __shfl_sync(0xffffffffu, po, lane | 1);
            if (!live) continue;
            const float coef = rsqrt_approx_ftz(sumsq / 128.f + 1e-6f);
            const float *w = normw + (part & 1) * 128 + sub * 8;
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                T2 hv = *reinterpret_cast<const T2 *>(&xw[e]);
                if (part < 2) {
                    const float x0 = r[e].x * (coef * w[2 * e]);
                    const float x1 = r[e].y * (coef * w[2 * e + 1]);
                    float y0 = x0 * sc[e].y - x1 * sc[e].x;
                    float y1 = x0 * sc[e].x + x1 * sc[e].y;
                    if constexpr (!Tr::kIsBf16) {   // fp16 stores clamp (gemm_base.cuh:688-696)
                        y0 = fminf(fmaxf(y0, -65504.f), 65504.f);
                        y1 = fminf(fmaxf(y1, -65504.f), 65504.f);
                    }
                    hv = Tr::from_float2(make_float2(y0, y1));
                }
                if constexpr (PACK) {
                    const __half2 h2 = __float22half2_rn(Tr::to_float2(hv));
                    o[e] = *reinterpret_cast<const uint32_t *>(&h2);
                } else {
                    o[e] = *reinterpret_cast<const uint32_t *>(&hv);
                }
            }
            if constexpr (PACK) {
                __half *dst = (part == 0 ? out_q + hh * sq : part == 1 ? out_k + hh * sk : out_v + hh * sv) + static_cast<size_t>(m) * 128 + sub * 8;
                *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
                *reinterpret_cast<uint4 *>(row + static_cast<size_t>(head) * 128) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

// NB200_ROPE = 1 | 2 selects the lane mapping (read at every launch: tools/rope_bench.py times both in one process)
inline int rope_variant() {
    const char *e = getenv("NB200_ROPE");
    return e != nullptr && (e[0] == '1' || e[0] == '2') ? e[0] - '0' : kRopeDefaultVariant;
}

}  // namespace

// qkv hT [M, N] (N = 3 * heads * 128), rotary: the reference's packed table covering at least the rows < M
int rope_inplace_dispatch(int dtype, void *qkv, int M, int N, const void *norm_q, const void *norm_k, const float *rotary, cudaStream_t stream) {
    NB200_REQUIRE(qkv && norm_q && norm_k && rotary, "NULL tensor");
    NB200_REQUIRE(N % 384 == 0, "N must be 3 * heads * 128");
    if (M <= 0) return NB200_OK;
    const unsigned grid = static_cast<unsigned>((M + kRopeWarps - 1) / kRopeWarps);
    LaunchCfg lc(dim3(grid), dim3(kRopeWarps * 32), 0, stream);
    __half *none = nullptr;
    const bool v2 = rope_variant() == 2;
    if (dtype == NB200_BF16) {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, v2 ? rope16_kernel<__nv_bfloat16, false> : rope_kernel<__nv_bfloat16, false>, static_cast<__nv_bfloat16 *>(qkv), M, N,
                                            static_cast<const __nv_bfloat16 *>(norm_q), static_cast<const __nv_bfloat16 *>(norm_k), rotary, none, none, none, 0ll, 0ll,
                                            0ll, M));
    } else {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, v2 ? rope16_kernel<__half, false> : rope_kernel<__half, false>, static_cast<__half *>(qkv), M, N, static_cast<const __half *>(norm_q),
                                            static_cast<const __half *>(norm_k), rotary, none, none, none, 0ll, 0ll, 0ll, M));
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
    const unsigned grid = static_cast<unsigned>((Mp + kRopeWarps - 1) / kRopeWarps);
    LaunchCfg lc(dim3(grid), dim3(kRopeWarps * 32), 0, stream);
    const bool v2 = rope_variant() == 2;
    if (dtype == NB200_BF16) {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, v2 ? rope16_kernel<__nv_bfloat16, true> : rope_kernel<__nv_bfloat16, true>, static_cast<__nv_bfloat16 *>(const_cast<void *>(qkv)), Mp, N,
                                            static_cast<const __nv_bfloat16 *>(norm_q), static_cast<const __nv_bfloat16 *>(norm_k), rotary,
                                            static_cast<__half *>(out_q), static_cast<__half *>(out_k), static_cast<__half *>(out_v), sq, sk, sv, attn_tokens));
    } else {
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, v2 ? rope16_kernel<__half, true> : rope_kernel<__half, true>, static_cast<__half *>(const_cast<void *>(qkv)), Mp, N, static_cast<const __half *>(norm_q),
                                            static_cast<const __half *>(norm_k), rotary, static_cast<__half *>(out_q), static_cast<__half *>(out_k),
                                            static_cast<__half *>(out_v), sq, sk, sv, attn_tokens));
    }
    count_launch();
    return NB200_OK;
}

}  // namespace nb200
