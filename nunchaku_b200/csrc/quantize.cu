// Activation quantiser + low-rank down projection for the SVDQuant W4A4 linear (B200).
//
// Replaces quantize_w4a4_fuse_lora_kernel (reference src/kernels/zgemm/gemm_w4a4.cuh:1097-1184,
// launcher gemm_w4a4_launch_impl.cuh:451-521).  HBM-bound: reads x once (2*M*K bytes), writes
// M*K/2 + scales + 4*M*R bytes.  Design:
//   * one CTA owns 16*MT complete rows, so the rank-R projection is reduced inside the CTA in a
//     fixed order -- no atomics, bit-reproducible (the reference uses red.global.add.f32);
//   * every lane streams 16-byte pieces of its two rows straight into registers; the same
//     registers feed (a) mma.sync.m16n8k16 for the skinny x @ lora_down^T (k-permuted fragments:
//     the contraction is order-free, so "slot 2t/2t+8" are bound to 4 consecutive k and
//     lora_down is pre-shuffled to match at load time) and (b) the per-group absmax / 4-bit
//     rounding, which uses the reference's exact instruction recipe (div.approx, rcp.approx.ftz,
//     cvt.rni + saturating pack, e2m1/e4m3 cvt) so codes match bit for bit where the inputs do;
//   * 8 warps split K by 64-wide groups; per-warp fp32 partials are summed through shared memory.
//
// Output layouts are the B200 inter-op layouts of include/nunchaku_b200.h.
#include "common.cuh"

namespace nb200 {
namespace {

struct QParams {
    const void *x;
    uint8_t *q;
    void *scales;
    const void *ld;  // [K/32][Rp/8][32][8] hT
    float *lora;     // [Mp, R]
    const void *smooth;
    int M, Mp, K, R, Rp;
    int x_stride;  // elements per input row (K, or 2K with GLU)
};

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

template <typename hT>
__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                             uint32_t b0, uint32_t b1) {
    if constexpr (HalfTraits<hT>::kIsBf16) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
            "{%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    } else {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
            "{%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
}

template <typename hT>
__device__ __forceinline__ typename HalfTraits<hT>::T2 as_h2(uint32_t v) {
    return *reinterpret_cast<typename HalfTraits<hT>::T2 *>(&v);
}
template <typename hT>
__device__ __forceinline__ uint32_t as_u32(typename HalfTraits<hT>::T2 v) {
    return *reinterpret_cast<uint32_t *>(&v);
}

// 8 consecutive (post-GLU) activations of one row as 4 packed words; zeros beyond M.
template <typename hT, bool GLU>
__device__ __forceinline__ uint4 load_x8(const hT *row, int k, bool valid) {
    using Tr = HalfTraits<hT>;
    if (!valid) return make_uint4(0, 0, 0, 0);
    if constexpr (!GLU) {
        return ldg_nc_v4(row + k);
    } else {
        // x'[j] = x[2j] * silu(x[2j+1]);  silu and product rounded to hT (gemm_base.cuh:612-623)
        const uint4 lo = ldg_nc_v4(row + 2 * k);
        const uint4 hi = ldg_nc_v4(row + 2 * k + 8);
        const uint32_t in[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        uint32_t out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 p0 = Tr::to_float2(as_h2<hT>(in[2 * i]));
            const float2 p1 = Tr::to_float2(as_h2<hT>(in[2 * i + 1]));
            const hT s0 = Tr::from_float(silu_f32(p0.y));
            const hT s1 = Tr::from_float(silu_f32(p1.y));
            typename Tr::T2 r;
            r.x = __hmul(Tr::from_float(p0.x), s0);
            r.y = __hmul(Tr::from_float(p1.x), s1);
            out[i] = as_u32<hT>(r);
        }
        return make_uint4(out[0], out[1], out[2], out[3]);
    }
}

template <typename hT>
__device__ __forceinline__ void smooth8(const uint4 &x, const uint4 &sm, float (&xs)[8], float &amax) {
    using Tr = HalfTraits<hT>;
    const uint32_t xv[4] = {x.x, x.y, x.z, x.w};
    const uint32_t sv[4] = {sm.x, sm.y, sm.z, sm.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 a = Tr::to_float2(as_h2<hT>(xv[i]));
        const float2 b = Tr::to_float2(as_h2<hT>(sv[i]));
        // h2div: fp32 __fdividef, rounded back to hT (gemm_utils.cuh:329-344)
        const float2 d = Tr::to_float2(Tr::from_float2(make_float2(__fdividef(a.x, b.x), __fdividef(a.y, b.y))));
        xs[2 * i] = d.x;
        xs[2 * i + 1] = d.y;
        amax = fmaxf(amax, fmaxf(fabsf(d.x), fabsf(d.y)));  // max of hT values is exact in fp32
    }
}

template <typename hT>
__device__ __forceinline__ void plain8(const uint4 &x, float (&xs)[8], float &amax) {
    using Tr = HalfTraits<hT>;
    const uint32_t xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 a = Tr::to_float2(as_h2<hT>(xv[i]));
        xs[2 * i] = a.x;
        xs[2 * i + 1] = a.y;
        amax = fmaxf(amax, fmaxf(fabsf(a.x), fabsf(a.y)));
    }
}

template <typename hT, bool FP4, bool GLU, int MT>
__global__ void __launch_bounds__(kThreads) quantize_kernel(const QParams p) {
    using Tr = HalfTraits<hT>;
    constexpr int ROWS = 16 * MT;
    __shared__ float red[kWarps][ROWS][33];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int gq = lane >> 2;  // row within 8
    const int t = lane & 3;    // 8-element piece within a 32-wide k block
    const int row_base = blockIdx.x * ROWS;
    const int num_groups = p.K >> 6;
    const hT *x = reinterpret_cast<const hT *>(p.x);
    const hT *smooth = reinterpret_cast<const hT *>(p.smooth);
    const uint4 *ldw = reinterpret_cast<const uint4 *>(p.ld);
    const int nt_total = p.Rp >> 3;  // 8-rank tiles in lora_down

    for (int chunk = 0; chunk * 32 < p.Rp; chunk++) {
        float acc[MT][4][4];
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[m][j][e] = 0.f;

        for (int g = warp; g < num_groups; g += kWarps) {
            // ---- loads: [mt][kb][row 0/1] 16 bytes each --------------------------------------
            uint4 xa[MT][2][2];
#pragma unroll
            for (int m = 0; m < MT; m++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int r = row_base + m * 16 + gq + h * 8;
                    const bool valid = r < p.M;
                    const hT *rowp = x + static_cast<size_t>(r) * p.x_stride;
#pragma unroll
                    for (int kb = 0; kb < 2; kb++) xa[m][kb][h] = load_x8<hT, GLU>(rowp, g * 64 + kb * 32 + t * 8, valid);
                }
            }
            uint4 bw[2][4];
#pragma unroll
            for (int kb = 0; kb < 2; kb++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    bw[kb][j] = ldg_v4(ldw + (static_cast<size_t>(g * 2 + kb) * nt_total + chunk * 4 + j) * 32 + lane);

            // ---- low-rank down projection on the un-smoothed tile (lora.cuh:243-353) -----------
#pragma unroll
            for (int m = 0; m < MT; m++) {
#pragma unroll
                for (int kb = 0; kb < 2; kb++) {
                    const uint32_t a_lo[4] = {xa[m][kb][0].x, xa[m][kb][0].y, xa[m][kb][0].z, xa[m][kb][0].w};
                    const uint32_t a_hi[4] = {xa[m][kb][1].x, xa[m][kb][1].y, xa[m][kb][1].z, xa[m][kb][1].w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t b[4] = {bw[kb][j].x, bw[kb][j].y, bw[kb][j].z, bw[kb][j].w};
#pragma unroll
                        for (int s = 0; s < 2; s++)
                            mma_m16n8k16<hT>(acc[m][j], a_lo[2 * s], a_hi[2 * s], a_lo[2 * s + 1], a_hi[2 * s + 1],
                                             b[2 * s], b[2 * s + 1]);
                    }
                }
            }

            if (chunk != 0) continue;

            // ---- smooth + per-group quantisation (gemm_w4a4.cuh:85-187, 429-523, 961-1003) ------
            uint4 sm[2];
            if (smooth != nullptr) {
#pragma unroll
                for (int kb = 0; kb < 2; kb++) sm[kb] = ldg_v4(smooth + g * 64 + kb * 32 + t * 8);
            }
#pragma unroll
            for (int m = 0; m < MT; m++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int r = row_base + m * 16 + gq + h * 8;  // < Mp always
                    float xs[2][8];
                    float amax_kb[2] = {0.f, 0.f};
#pragma unroll
                    for (int kb = 0; kb < 2; kb++) {
                        if (smooth != nullptr)
                            smooth8<hT>(xa[m][kb][h], sm[kb], xs[kb], amax_kb[kb]);
                        else
                            plain8<hT>(xa[m][kb][h], xs[kb], amax_kb[kb]);
                    }
                    uint8_t *qrow = p.q + static_cast<size_t>(r) * (p.K >> 1);
                    if constexpr (!FP4) {
                        float amax = fmaxf(amax_kb[0], amax_kb[1]);
                        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
                        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
                        const float s32 = amax * (1.0f / 7.0f);
                        const float rs = rcp_approx_ftz(s32);
                        if (t == 0)
                            reinterpret_cast<hT *>(p.scales)[static_cast<size_t>(g) * p.Mp + r] = Tr::from_float(s32);
#pragma unroll
                        for (int kb = 0; kb < 2; kb++) {
                            int qv[8];
#pragma unroll
                            for (int e = 0; e < 8; e++) qv[e] = cvt_rni(xs[kb][e] * rs);
                            const uint32_t w = pack8_int4_b200<false>(qv);
                            *reinterpret_cast<uint32_t *>(qrow + ((g * 64 + kb * 32 + t * 8) >> 1)) = w;
                        }
                    } else {
                        uint32_t sbyte[2];
#pragma unroll
                        for (int kb = 0; kb < 2; kb++) {
                            float amax = fmaxf(amax_kb[kb], __shfl_xor_sync(0xffffffffu, amax_kb[kb], 1));
                            const float sc = fminf(amax * (1.0f / 6.0f), 448.0f);
                            const float rs = rcp_approx_ftz(sc);
                            sbyte[kb] = cvt_e4m3x2(0.f, sc) & 0xFFu;
                            uint32_t w = 0;
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                w |= cvt_e2m1x2(xs[kb][2 * i + 1] * rs, xs[kb][2 * i] * rs) << (8 * i);
                            *reinterpret_cast<uint32_t *>(qrow + ((g * 64 + kb * 32 + t * 8) >> 1)) = w;
                        }
                        // 16-groups of this 64 block: c = kb*2 + t/2; gather the four bytes in lane t == 0
                        const uint32_t mine = sbyte[0] | (sbyte[1] << 16);           // c0 (or c1), c2 (or c3)
                        const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 2);  // lanes t^2
                        if (t == 0) {
                            const uint32_t word = mine | (other << 8);  // c0 | c1<<8 | c2<<16 | c3<<24
                            uint8_t *sf = reinterpret_cast<uint8_t *>(p.scales) +
                                          (static_cast<size_t>(r >> 7) * num_groups + g) * 512 + (r & 31) * 16 +
                                          ((r & 127) >> 5) * 4;
                            *reinterpret_cast<uint32_t *>(sf) = word;
                        }
                    }
                }
            }
        }

        // ---- deterministic cross-warp reduction of the low-rank partials -----------------------
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                red[warp][m * 16 + gq][j * 8 + t * 2] = acc[m][j][0];
                red[warp][m * 16 + gq][j * 8 + t * 2 + 1] = acc[m][j][1];
                red[warp][m * 16 + gq + 8][j * 8 + t * 2] = acc[m][j][2];
                red[warp][m * 16 + gq + 8][j * 8 + t * 2 + 1] = acc[m][j][3];
            }
        __syncthreads();
        for (int i = threadIdx.x; i < ROWS * 32; i += kThreads) {
            const int rr = i >> 5, c = i & 31;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; w++) s += red[w][rr][c];
            const int rank = chunk * 32 + c;
            if (rank < p.R) p.lora[static_cast<size_t>(row_base + rr) * p.R + rank] = s;
        }
        __syncthreads();
    }
}

template <typename hT, bool FP4, bool GLU>
int launch(const QParams &p, cudaStream_t stream) {
    // 32-row CTAs halve the L2 traffic of lora_down; use 16-row CTAs when that would leave
    // most of the 148 SMs idle.
    const bool big = (p.Mp / 32) >= 120;
    if (big) {
        quantize_kernel<hT, FP4, GLU, 2><<<p.Mp / 32, kThreads, 0, stream>>>(p);
    } else {
        quantize_kernel<hT, FP4, GLU, 1><<<p.Mp / 16, kThreads, 0, stream>>>(p);
    }
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    return NB200_OK;
}

}  // namespace

int quantize_v2_dispatch(const nb200_quantize_args &a, cudaStream_t stream);

}  // namespace nb200

extern "C" __attribute__((visibility("default"))) long long nb200_quantize_workspace_bytes(int Mp, int K) {
    if (Mp <= 0 || K <= 0) return 0;
    const long long row_blocks = Mp / 32;
    // K split so that the row blocks x splits fill 2 CTA slots on each of up to 160 SMs, never finer than 256 k per split (quantize_v2.cu)
    long long ks_max = (2 * 160 + row_blocks - 1) / row_blocks;
    if (ks_max > K / 256) ks_max = K / 256;
    if (ks_max < 1) ks_max = 1;
    return ((row_blocks * 4 + 255) / 256) * 256 + ks_max * Mp * 32 * 4;
}

extern "C" __attribute__((visibility("default"))) int nb200_quantize_w4a4_act_fuse_lora(const nb200_quantize_args *a, void *stream_) {
    using namespace nb200;
    reset_launch_count();
    NB200_REQUIRE(a != nullptr, "args is NULL");
    NB200_REQUIRE(a->input && a->output && a->oscales, "NULL tensor");
    NB200_REQUIRE(a->R == 0 || (a->lora_down && a->lora_act_out), "lora_down and lora_act_out are required when rank > 0");
    NB200_REQUIRE(a->M > 0 && a->M <= a->Mp, "M must be in (0, Mp]");
    NB200_REQUIRE(a->Mp % 256 == 0, "Mp must be a multiple of 256 (pad_size)");
    NB200_REQUIRE(a->K % 128 == 0 && a->K > 0, "K must be a positive multiple of 128");
    NB200_REQUIRE(a->R % 16 == 0 && a->R >= 0, "rank must be a multiple of 16 (0: quantise only, as the reference's rank-0 GEMM_W4A4)");
    NB200_REQUIRE(a->R > 0 || !a->fuse_glu, "rank 0 is not supported together with fuse_glu");
    NB200_REQUIRE(a->dtype == NB200_FP16 || a->dtype == NB200_BF16, "dtype must be fp16 or bf16");
    NB200_REQUIRE((reinterpret_cast<uintptr_t>(a->input) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->output) & 15) == 0,
                  "input/output must be 16-byte aligned");
    NB200_REQUIRE(!a->act_unsigned_shift || (!a->fp4 && !a->fuse_glu), "act_unsigned_shift is INT4 only and excludes fuse_glu");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!a->fuse_glu) return quantize_v2_dispatch(*a, stream);
    QParams p;
    p.x = a->input;
    p.q = static_cast<uint8_t *>(a->output);
    p.scales = a->oscales;
    // second half of the repacked factor: k-permuted fragments (nb200_repack_lora_down)
    p.ld = static_cast<const uint8_t *>(a->lora_down) + static_cast<size_t>(a->K) * ((a->R + 31) / 32 * 32) * 2;
    p.lora = a->lora_act_out;
    p.smooth = a->smooth;
    p.M = a->M;
    p.Mp = a->Mp;
    p.K = a->K;
    p.R = a->R;
    p.Rp = (a->R + 31) / 32 * 32;
    p.x_stride = a->fuse_glu ? 2 * a->K : a->K;
    const bool bf16 = a->dtype == NB200_BF16;
#define NB200_Q_DISPATCH(HT)                                                         \
    if (a->fp4) {                                                                    \
        return a->fuse_glu ? launch<HT, true, true>(p, stream) : launch<HT, true, false>(p, stream);   \
    } else {                                                                         \
        return a->fuse_glu ? launch<HT, false, true>(p, stream) : launch<HT, false, false>(p, stream); \
    }
    if (bf16) {
        NB200_Q_DISPATCH(__nv_bfloat16)
    } else {
        NB200_Q_DISPATCH(__half)
    }
#undef NB200_Q_DISPATCH
}
