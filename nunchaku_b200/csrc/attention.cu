// Scaled-dot-product attention of the FLUX blocks on tcgen05.
//
// Replaces nunchaku::kernels::attention_fp16 (src/kernels/zgemm/attention.cu:10-94, kernel attention.cuh:420-703; Python
// NunchakuFP16AttnProcessor, nunchaku/models/attention_processors/flux.py): q / k / v fp16 [B, H, T, 128] as the QKV projection's PackQKV
// epilogue writes them (row-major inside a head: OUR layout -- the reference stores its own mma.sync fragment order there), pad rows of K
// are NaN and act as the mask (scores NaN -> -inf, attention.cuh:192-221), o hT [B, Tq, H * 128], softmax in base 2 with
// scale * log2(e) folded in (attention.cu:48-49), non-causal.
//
// One CTA per (128 query rows, head).  Flash-attention forward with both GEMMs on the 5th-generation tensor cores:
//   S = Q K^T   : tcgen05.mma kind::f16, M = 128, N = 128 keys, K = 128 (8 instructions); Q and K tiles are K-major SW128 TMA boxes;
//                 the S accumulator is double buffered in TMEM so that S(j+1) runs under the softmax of tile j
//   softmax     : 128 threads, thread = query row = TMEM lane (no shuffles): tcgen05.ld of the row's 128 scores, running max / sum in
//                 fp32, ex2.approx, P rounded to fp16 into shared memory in the K-major SW128 operand layout
//   O += P V    : A = P (shared memory), B = V as an MN-major SW128 operand (V is [keys, d] row-major: keys are the K dimension, d
//                 is contiguous -- no transpose anywhere), fp32 accumulator in TMEM; when a row's running max moved, O is rescaled in
//                 TMEM (tcgen05.ld / st) before the next P V is issued
//   epilogue    : O * rcp(l) -> hT -> global
// The reference accumulates Q K^T and P V in fp16 per 32-key tile (mma.sync f16 accumulate, attention.cuh:185-260,341-366); fp32
// accumulation here is strictly more accurate, so parity is to fp16 noise (tests/test_gpu_attention.py: vs fp64 softmax attention, vs
// the reference kernel live on the same GPU).
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace nb200 {
namespace {

using namespace ptx;

constexpr int kD = 128;            // head dimension
constexpr int kBM = 128;           // query rows per CTA
constexpr int kBN = 128;           // keys per tile
constexpr int kSlab = kBM * 128;   // one [128 rows x 128 B] SW128 box = 64 fp16 columns
constexpr int kTile = 2 * kSlab;   // a 128 x 128 fp16 tile: 32 KB
constexpr int kKvStages = 2;
constexpr int kThreadsAttn = 256;  // warps 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 softmax / correction / epilogue

struct alignas(1024) SmemA {
    alignas(1024) uint8_t q[kTile];
    alignas(1024) uint8_t k[kKvStages][kTile];
    alignas(1024) uint8_t v[kKvStages][kTile];
    alignas(1024) uint8_t p[kTile];
    uint64_t q_full;
    uint64_t k_full[kKvStages], k_empty[kKvStages];
    uint64_t v_full[kKvStages], v_empty[kKvStages];
    uint64_t s_full[2], s_empty[2];
    uint64_t p_full, pv_done;
    uint32_t tmem_base;
};
static_assert(sizeof(SmemA) + 1024 <= 232448, "shared memory budget");

struct AttnParams {
    void *o;
    int heads, tokens_q, tokens_kv;
    float scale_log2;   // scale * log2(e)
    int out_bf16;
    int debug;          // NB200_ATTN_DEBUG ablation bits (v2): 2 = always the two-pass path; results invalid: 4 = no exponentials, 8 = K / V loaded once, 16 / 32 = one eighth of the P V / Q K^T MMAs
};

// MN-major SW128 operand (V: [keys][64 d] boxes, two boxes for d = 128): start address + 16 keys per instruction;
// LBO = distance between the two 64-wide d blocks, SBO = distance between groups of 8 keys
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr) { return make_smem_desc(smem_addr, kSlab, 1024, kLayoutSw128); }

__global__ void __launch_bounds__(kThreadsAttn, 1)
attention_fp16_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                      const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    SmemA &s = *reinterpret_cast<SmemA *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int n_tiles = p.tokens_kv / kBN;
    const int q_row0 = (batch * p.heads + head) * p.tokens_q + qb * kBM;    // row in the [B * H * Tq, 128] view
    const int kv_row0 = (batch * p.heads + head) * p.tokens_kv;

    if (warp == 0 && elect_one()) {
        prefetch_tensormap(&tm_q);
        prefetch_tensormap(&tm_k);
        prefetch_tensormap(&tm_v);
        mbar_init(&s.q_full, 1);
        for (int i = 0; i < kKvStages; i++) {
            mbar_init(&s.k_full[i], 1);
            mbar_init(&s.k_empty[i], 1);
            mbar_init(&s.v_full[i], 1);
            mbar_init(&s.v_empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&s.s_full[i], 1);
            mbar_init(&s.s_empty[i], 4);
        }
        mbar_init(&s.p_full, 4);
        mbar_init(&s.pv_done, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(&s.tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s.tmem_base;
    const uint32_t tmem_s[2] = {tmem_base, tmem_base + 128};
    const uint32_t tmem_o = tmem_base + 256;
    griddep_launch_dependents();
    griddep_wait();

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            mbar_expect_tx(&s.q_full, kTile);
            tma_load_2d(s.q, &tm_q, &s.q_full, 0, q_row0);
            tma_load_2d(s.q + kSlab, &tm_q, &s.q_full, 64, q_row0);
            for (int j = 0; j < n_tiles; j++) {
                const int st = j % kKvStages;
                const uint32_t ph = (j / kKvStages) & 1;
                mbar_wait(&s.k_empty[st], ph ^ 1);
                mbar_expect_tx(&s.k_full[st], kTile);
                tma_load_2d(s.k[st], &tm_k, &s.k_full[st], 0, kv_row0 + j * kBN);
                tma_load_2d(s.k[st] + kSlab, &tm_k, &s.k_full[st], 64, kv_row0 + j * kBN);
                mbar_wait(&s.v_empty[st], ph ^ 1);
                mbar_expect_tx(&s.v_full[st], kTile);
                tma_load_2d(s.v[st], &tm_v, &s.v_full[st], 0, kv_row0 + j * kBN);
                tma_load_2d(s.v[st] + kSlab, &tm_v, &s.v_full[st], 64, kv_row0 + j * kBN);
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        if (elect_one()) {
            constexpr uint32_t idesc_qk = make_idesc_f16(false, kBM, kBN);
            constexpr uint32_t idesc_pv = make_idesc_f16(false, kBM, kD) | (1u << 16);   // B (= V) is MN-major
            const uint32_t q_addr = smem_u32(s.q), p_addr = smem_u32(s.p);
            auto issue_s = [&](int j) {   // S(j) = Q K_j^T into S buffer j % 2
                const int st = j % kKvStages, b = j & 1;
                mbar_wait(&s.k_full[st], (j / kKvStages) & 1);
                mbar_wait(&s.s_empty[b], ((j >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t k_addr = smem_u32(s.k[st]);
#pragma unroll
                for (int ks = 0; ks < kD / 16; ks++) {
                    const uint32_t off = (ks >> 2) * kSlab + (ks & 3) * 32;
                    tc_mma_f16(tmem_s[b], make_sw128_kmajor_desc(q_addr + off), make_sw128_kmajor_desc(k_addr + off), idesc_qk, ks != 0);
                }
                tc_commit(&s.k_empty[st]);
                tc_commit(&s.s_full[b]);
            };
            mbar_wait(&s.q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_tiles; j++) {
                if (j + 1 < n_tiles) issue_s(j + 1);
                const int st = j % kKvStages;
                mbar_wait(&s.p_full, j & 1);
                mbar_wait(&s.v_full[st], (j / kKvStages) & 1);
                tc_fence_after_sync();
                const uint32_t v_addr = smem_u32(s.v[st]);
#pragma unroll
                for (int ks = 0; ks < kBN / 16; ks++) {
                    const uint32_t a_off = (ks >> 2) * kSlab + (ks & 3) * 32;   // P: K-major, keys contiguous
                    const uint32_t b_off = ks * 16 * 128;                       // V: 16 key rows of 128 B further
                    tc_mma_f16(tmem_o, make_sw128_kmajor_desc(p_addr + a_off), make_sw128_mnmajor_desc(v_addr + b_off), idesc_pv, (j | ks) != 0);
                }
                tc_commit(&s.v_empty[st]);
                tc_commit(&s.pv_done);
            }
        }
    } else if (warp >= 4) {
        // ============================ softmax / correction / epilogue: thread = query row ==============================
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(qd * 32) << 16;
        const float scale = p.scale_log2;
        float m_run = -INFINITY, l_run = 0.f;
        const uint32_t p_row = smem_u32(s.p) + row * 128;
        for (int j = 0; j < n_tiles; j++) {
            const int b = j & 1;
            mbar_wait(&s.s_full[b], (j >> 1) & 1);
            tc_fence_after_sync();
            uint32_t sv[4][32];
#pragma unroll
            for (int c = 0; c < 4; c++) tmem_ld_32x32b_x32(tmem_s[b] + lane_base + c * 32, sv[c]);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.s_empty[b]);
            // running max (NaN scores = masked keys: fmaxf drops the NaN operand, attention.cuh:192-221)
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int i = 0; i < 32; i++) mx = fmaxf(mx, __uint_as_float(sv[c][i]));
            const float m_new = fmaxf(m_run, mx * scale);
            const float alpha = ex2_approx_ftz(m_run - m_new);   // first tile: ex2(-inf) = 0
            // p = 2^(s * scale - m_new); masked (NaN) scores -> 0
            float sum = 0.f;
            uint32_t pk[4][16];
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float s0 = fmaxf(__uint_as_float(sv[c][i]), -INFINITY), s1 = fmaxf(__uint_as_float(sv[c][i + 1]), -INFINITY);
                    const float e0 = ex2_approx_ftz(fmaf(s0, scale, -m_new)), e1 = ex2_approx_ftz(fmaf(s1, scale, -m_new));
                    sum += e0 + e1;
                    const __half2 h = __floats2half2_rn(e0, e1);
                    pk[c][i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
                }
            l_run = fmaf(l_run, alpha, sum);
            m_run = m_new;
            // P V of the previous tile must be done before P is overwritten and before O is touched
            if (j > 0) {
                mbar_wait(&s.pv_done, (j - 1) & 1);
                tc_fence_after_sync();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {   // some row of this warp moved its maximum: rescale the warp's 32 rows of O
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t ov[32];
                        tmem_ld_32x32b_x32(tmem_o + lane_base + c * 32, ov);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i++) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                        tmem_st_32x32b_x32(tmem_o + lane_base + c * 32, ov);
                    }
                    tmem_st_wait();
                }
            }
            // P -> shared memory: [128 rows][2 slabs of 64 keys], 16-byte chunks XOR-swizzled by the row (SW128 K-major operand layout)
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int chunk = (c & 1) * 4 + u;   // 16-byte chunk inside the slab's 128-byte row
                    const uint32_t addr = p_row + (c >> 1) * kSlab + ((chunk ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[c][4 * u]), "r"(pk[c][4 * u + 1]), "r"(pk[c][4 * u + 2]),
                                 "r"(pk[c][4 * u + 3])
                                 : "memory");
                }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.p_full);
        }
        // ---- epilogue: O / l -> hT -> o[batch, q, head * 128 + d]
        mbar_wait(&s.pv_done, (n_tiles - 1) & 1);
        tc_fence_after_sync();
        const float inv = rcp_approx_ftz(l_run);
        const size_t orow = (static_cast<size_t>(batch) * p.tokens_q + qb * kBM + row) * (static_cast<size_t>(p.heads) * kD) + static_cast<size_t>(head) * kD;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(tmem_o + lane_base + c * 32, ov);
            tmem_ld_wait();
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float a = __uint_as_float(ov[i]) * inv, bb = __uint_as_float(ov[i + 1]) * inv;
                if (p.out_bf16) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(a, bb);
                    w[i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
                } else {
                    const __half2 h = __floats2half2_rn(a, bb);
                    w[i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
                }
            }
            uint4 *dst = reinterpret_cast<uint4 *>(static_cast<uint16_t *>(p.o) + orow + c * 32);
#pragma unroll
            for (int u = 0; u < 4; u++) dst[u] = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
        }
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}


// ------------------------------------------------------------------------------------------------------------------------------
// v2 (default): TWO CTAs per SM, P handed to the tensor core through TMEM.
//
// What bounded v1 (tools/attn_bench.py, r02: 340 us at 24 heads x 4608 tokens = 34 % of the 16-bit tensor peak): one CTA per SM whose
// 128 softmax threads (one warp per scheduler, no latency hiding) sat between the two GEMMs of every key tile -- tcgen05.ld, 128 ex2,
// fp16 pack, 16 st.shared, a proxy fence, the P V MMAs, then the next tile: ~2.9k clk per tile against 1.0k clk of MMA time.
// v2 keeps the thread = query row softmax and changes what surrounds it:
//   * P never touches shared memory: the softmax threads write the fp16 probabilities back into TMEM over the first 64 columns of S
//     (tcgen05.st, two keys per 32-bit column) and O += P V is issued with the A operand read from TMEM
//     (tcgen05.mma [d], [a_tmem], b_desc).  No 32 KB P buffer, no st.shared, no fence.proxy.async;
//   * shared memory per CTA drops to Q + one K tile + one V tile = 96 KB and TMEM to 256 columns (S|P 128 + O 128), so TWO CTAs are
//     resident per SM: while one CTA's threads are in their softmax the other CTA's MMAs own the tensor pipe (and its MUFU work
//     interleaves on the same schedulers);
//   * S is read twice from TMEM (running maximum, then exponentials 32 columns at a time) instead of being held in 128 registers;
//   * the running maximum is only moved -- and O only rescaled -- when a tile raises it by more than 2^8 (probabilities stay <= 256, far
//     inside fp16; the final division by the running sum makes the result independent of the stale maximum);
//   * keys masked by NaN (the reference's convention, attention.cuh:192-221) are detected with one NaN-propagating max per element in
//     the maximum pass; only tiles that contain masked keys take the path that cleans every score.
// Measured (B200, 24 heads x 4608 tokens, tools/attn_bench.py): v1 328-340 us, v2 221-230 us (the reference kernel built for sm_100a 1066 us, the
// library SDPA 156-170 us).  What this shape is bound by (ablations, NB200_ATTN_DEBUG): not the tensor pipe (7/8 of the MMAs removed: 231 -> 189 us)
// and not MUFU (exponentials removed: no change) but the serial chain of one CTA -- commit -> wake-up -> tcgen05.ld round trips -> ex2 ->
// tcgen05.st -> arrive -> P V -> Q K'^T -- of which two run per SM.  Tried on top of it and measured no faster (kept as
// profiles/r02_attention_pingpong_experiments.patch with the timeline tool that produced the numbers): one CTA per SM with two query tiles in
// FlashAttention-4's ping-pong order and S held in 128 registers (252 us); the same with two softmax threads per row (244 us); 64-key tiles with
// a double-buffered S per chain (272 us: tools/ubench/mma_shapes.cu -- an N = 64 instruction costs 58 clk, N = 128 75 clk, the A-from-TMEM
// form 79 clk against 32 / 64 / 64 nominal, so small tiles pay per instruction); every fourth pair of exponentials as a degree-3 polynomial on
// the FMA pipes (no gain: inside one in-order warp MUFU and FMA-pipe time add up either way).
// In-order execution of the tensor pipe is what makes the S|P alias safe: S(j+1) = Q K^T is issued after P(j) V and overwrites columns the
// earlier instruction has finished reading.
constexpr int kThreadsV2 = 192;    // warp 0 TMA, warp 1 TMEM alloc + MMA, warps 2-5 softmax (TMEM lane quadrant = warp & 3)
constexpr float kRescaleThreshold = 8.f;

struct alignas(1024) SmemA2 {
    alignas(1024) uint8_t q[kTile];
    alignas(1024) uint8_t k[kTile];
    alignas(1024) uint8_t v[kTile];
    uint64_t q_full, k_full, k_empty, v_full, v_empty, s_full, p_full, o_full;
    uint32_t tmem_base;
};
static_assert(2 * (sizeof(SmemA2) + 1024 + 1024) <= 233472, "two CTAs per SM");

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 rows (lanes) x 16 k as 8 columns of packed 16-bit pairs
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ float max_nan_f32(float a, float b) {   // NaN-propagating max
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// (a.x * b.x + c.x, a.y * b.y + c.y) in one FFMA2 / (a + b) in one FADD2 (sm_100 packed fp32 pipe)
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<uint64_t &>(d))
        : "l"(reinterpret_cast<const uint64_t &>(a)), "l"(reinterpret_cast<const uint64_t &>(b)), "l"(reinterpret_cast<const uint64_t &>(c)));
    return d;
}
__device__ __forceinline__ float2 fadd2(const float2 a, const float2 b) {
    float2 d;
    asm("add.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t &>(d)) : "l"(reinterpret_cast<const uint64_t &>(a)), "l"(reinterpret_cast<const uint64_t &>(b)));
    return d;
}


__global__ void __launch_bounds__(kThreadsV2, 2)
attention_fp16_v2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                         const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    SmemA2 &s = *reinterpret_cast<SmemA2 *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int n_tiles = p.tokens_kv / kBN;
    const int q_row0 = (batch * p.heads + head) * p.tokens_q + qb * kBM;
    const int kv_row0 = (batch * p.heads + head) * p.tokens_kv;

    if (warp == 0 && elect_one()) {
        prefetch_tensormap(&tm_q);
        prefetch_tensormap(&tm_k);
        prefetch_tensormap(&tm_v);
        mbar_init(&s.q_full, 1);
        mbar_init(&s.k_full, 1);
        mbar_init(&s.k_empty, 1);
        mbar_init(&s.v_full, 1);
        mbar_init(&s.v_empty, 1);
        mbar_init(&s.s_full, 1);
        mbar_init(&s.p_full, 4);
        mbar_init(&s.o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<256>(&s.tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s.tmem_base;
    const uint32_t tmem_s = tmem_base, tmem_p = tmem_base, tmem_o = tmem_base + 128;
    griddep_launch_dependents();
    griddep_wait();

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            mbar_expect_tx(&s.q_full, kTile);
            tma_load_2d(s.q, &tm_q, &s.q_full, 0, q_row0);
            tma_load_2d(s.q + kSlab, &tm_q, &s.q_full, 64, q_row0);
            for (int j = 0; j < ((p.debug & 8) ? 1 : n_tiles); j++) {   // ablation 8: one K / V tile, reused (results invalid)
                const uint32_t ph = j & 1;
                mbar_wait(&s.k_empty, ph ^ 1);          // S(j-1) = Q K^T has retired
                mbar_expect_tx(&s.k_full, kTile);
                tma_load_2d(s.k, &tm_k, &s.k_full, 0, kv_row0 + j * kBN);
                tma_load_2d(s.k + kSlab, &tm_k, &s.k_full, 64, kv_row0 + j * kBN);
                mbar_wait(&s.v_empty, ph ^ 1);          // P(j-1) V has retired
                mbar_expect_tx(&s.v_full, kTile);
                tma_load_2d(s.v, &tm_v, &s.v_full, 0, kv_row0 + j * kBN);
                tma_load_2d(s.v + kSlab, &tm_v, &s.v_full, 64, kv_row0 + j * kBN);
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        if (elect_one()) {
            constexpr uint32_t idesc_qk = make_idesc_f16(false, kBM, kBN);
            constexpr uint32_t idesc_pv = make_idesc_f16(false, kBM, kD) | (1u << 16);   // B (= V) is MN-major
            const uint32_t q_addr = smem_u32(s.q), k_addr = smem_u32(s.k), v_addr = smem_u32(s.v);
            mbar_wait(&s.q_full, 0);
            for (int j = 0; j < n_tiles; j++) {
                const uint32_t ph = j & 1;
                if (j == 0 || !(p.debug & 8)) mbar_wait(&s.k_full, ph);
                tc_fence_after_sync();
#pragma unroll
                for (int ks = 0; ks < ((p.debug & 32) ? 1 : kD / 16); ks++) {   // ablation 32: one of the 8 Q K^T instructions
                    const uint32_t off = (ks >> 2) * kSlab + (ks & 3) * 32;
                    tc_mma_f16(tmem_s, make_sw128_kmajor_desc(q_addr + off), make_sw128_kmajor_desc(k_addr + off), idesc_qk, ks != 0);
                }
                tc_commit(&s.k_empty);
                tc_commit(&s.s_full);
                if (j == 0 || !(p.debug & 8)) mbar_wait(&s.v_full, ph);   // (long there: waited for first, nothing stands between P and its MMAs)
                mbar_wait(&s.p_full, ph);     // P(j) is in TMEM and O carries the current maximum
                tc_fence_after_sync();
#pragma unroll
                for (int ks = 0; ks < ((p.debug & 16) ? 1 : kBN / 16); ks++)   // ablation 16: one of the 8 P V instructions
                    tc_mma_f16_ts(tmem_o, tmem_p + ks * 8, make_sw128_mnmajor_desc(v_addr + ks * 16 * 128), idesc_pv, (j | ks) != 0);
                tc_commit(&s.v_empty);
            }
            tc_commit(&s.o_full);
        }
    } else {
        // ============================ softmax / correction / epilogue: thread = query row ==============================
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(qd * 32) << 16;
        const float scale = p.scale_log2;
        const float2 scale2 = make_float2(scale, scale);
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; j++) {
            mbar_wait(&s.s_full, j & 1);
            tc_fence_after_sync();
            // ---- fast path (every tile but the first, unless it raises the row maximum by more than 2^8 or holds masked keys): ONE pass over S
            //      with the maximum the row already has; the fp16 probabilities wait in registers until the whole row is known to be fine
            //      (P aliases S: nothing may be written before the decision)
            bool slow = j == 0 || (p.debug & 2);
            auto fast_pass = [&](auto mode_tag) {
                constexpr int kMode = decltype(mode_tag)::value;   // 1 = default, 4 = no exponentials (ablation)
                const float2 neg_m2 = make_float2(-m_run, -m_run);
                float2 sum2 = make_float2(0.f, 0.f);
                float tm0 = -INFINITY, tm1 = -INFINITY;
                uint32_t pk[64];
                uint32_t a[2][32];
                tmem_ld_32x32b_x32(tmem_s + lane_base, a[0]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (c < 3) tmem_ld_32x32b_x32(tmem_s + lane_base + (c + 1) * 32, a[(c + 1) & 1]);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float2 t = ffma2(make_float2(__uint_as_float(a[c & 1][i]), __uint_as_float(a[c & 1][i + 1])), scale2, neg_m2);
                        tm0 = max_nan_f32(tm0, t.x);
                        tm1 = max_nan_f32(tm1, t.y);
                        float2 e;
                        if constexpr (kMode == 4) e = t;
                        else e = make_float2(ex2_approx_ftz(t.x), ex2_approx_ftz(t.y));
                        sum2 = fadd2(sum2, e);
                        const __half2 h = __floats2half2_rn(e.x, e.y);
                        pk[c * 16 + (i >> 1)] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    if (c < 3) tmem_ld_wait();
                }
                const float tmax = max_nan_f32(tm0, tm1);
                slow = __any_sync(0xffffffffu, !(tmax <= kRescaleThreshold));   // (NaN compares false: masked keys take the slow path)
                if (!slow) {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t w[16];
#pragma unroll
                        for (int i = 0; i < 16; i++) w[i] = pk[c * 16 + i];
                        tmem_st_32x32b_x16(tmem_p + lane_base + c * 16, w);
                    }
                    l_run += sum2.x + sum2.y;
                }
            };
            if (!slow) {
                if (p.debug & 4) fast_pass(std::integral_constant<int, 4>{});
                else fast_pass(std::integral_constant<int, 1>{});
            }
            if (slow) {
                // ---- pass 1: maximum of the row's 128 scores; a NaN (masked key) poisons mx and sends the tile down the cleaning path
                float mx = -INFINITY;
    #pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    uint32_t a[32], b[32];
                    tmem_ld_32x32b_x32(tmem_s + lane_base + hf * 64, a);
                    tmem_ld_32x32b_x32(tmem_s + lane_base + hf * 64 + 32, b);
                    tmem_ld_wait();
                    float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    #pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        m0 = max_nan_f32(m0, __uint_as_float(a[i]));
                        m1 = max_nan_f32(m1, __uint_as_float(a[i + 1]));
                        m2 = max_nan_f32(m2, __uint_as_float(b[i]));
                        m3 = max_nan_f32(m3, __uint_as_float(b[i + 1]));
                    }
                    mx = max_nan_f32(max_nan_f32(m0, m1), max_nan_f32(m2, m3));
                }
                const bool masked = __any_sync(0xffffffffu, mx != mx);   // (all rows of a tile see the same masked keys)
                if (masked) {   // NaN-dropping maximum (fmaxf returns the other operand)
                    mx = -INFINITY;
    #pragma unroll 1
                    for (int c = 0; c < 4; c++) {
                        uint32_t a[32];
                        tmem_ld_32x32b_x32(tmem_s + lane_base + c * 32, a);
                        tmem_ld_wait();
    #pragma unroll
                        for (int i = 0; i < 32; i++) mx = fmaxf(mx, __uint_as_float(a[i]));
                    }
                }
                const float m_tile = mx * scale;
                float alpha = 1.f;
                if (m_tile > m_run + kRescaleThreshold) {   // first tile: m_run = -inf, alpha = 0
                    alpha = ex2_approx_ftz(m_run - m_tile);
                    m_run = m_tile;
                }
                const float neg_m = m_run == -INFINITY ? 0.f : -m_run;   // (a fully masked first tile: 2^(-inf - 0) = 0, not NaN)
                const float2 neg_m2 = make_float2(neg_m, neg_m);
                // ---- pass 2: p = 2^(s * scale - m_run), 32 scores at a time (the next 32 are in flight meanwhile); fp16 pairs go back into
                //      TMEM over columns S has vacated: chunk c of P = columns [16 c, 16 c + 16) <= the columns of S already consumed
                float2 sum2 = make_float2(0.f, 0.f);
                auto pass2 = [&](auto masked_tag) {
                    constexpr bool kMasked = decltype(masked_tag)::value;
                    uint32_t a[2][32];
                    tmem_ld_32x32b_x32(tmem_s + lane_base, a[0]);
                    tmem_ld_wait();
    #pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if (c < 3) tmem_ld_32x32b_x32(tmem_s + lane_base + (c + 1) * 32, a[(c + 1) & 1]);
                        uint32_t pk[16];
    #pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            float2 sv = make_float2(__uint_as_float(a[c & 1][i]), __uint_as_float(a[c & 1][i + 1]));
                            if constexpr (kMasked) {
                                sv.x = fmaxf(sv.x, -INFINITY);
                                sv.y = fmaxf(sv.y, -INFINITY);
                            }
                            const float2 t = ffma2(sv, scale2, neg_m2);
                            const float2 e = make_float2(ex2_approx_ftz(t.x), ex2_approx_ftz(t.y));
                            sum2 = fadd2(sum2, e);
                            const __half2 h = __floats2half2_rn(e.x, e.y);
                            pk[i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
                        }
                        tmem_st_32x32b_x16(tmem_p + lane_base + c * 16, pk);
                        if (c < 3) tmem_ld_wait();
                    }
                };
                if (masked) pass2(std::true_type{});
                else pass2(std::false_type{});
                l_run = fmaf(l_run, alpha, sum2.x + sum2.y);
                // ---- O carries the old maximum: rescale this warp's 32 rows when one of them moved (P(j-1) V retired before S(j) was signalled)
                if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
    #pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t ov[32];
                        tmem_ld_32x32b_x32(tmem_o + lane_base + c * 32, ov);
                        tmem_ld_wait();
    #pragma unroll
                        for (int i = 0; i < 32; i++) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                        tmem_st_32x32b_x32(tmem_o + lane_base + c * 32, ov);
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.p_full);
        }
        // ---- epilogue: O / l -> hT -> o[batch, q, head * 128 + d]
        mbar_wait(&s.o_full, 0);
        tc_fence_after_sync();
        const float inv = rcp_approx_ftz(l_run);
        const size_t orow = (static_cast<size_t>(batch) * p.tokens_q + qb * kBM + row) * (static_cast<size_t>(p.heads) * kD) + static_cast<size_t>(head) * kD;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(tmem_o + lane_base + c * 32, ov);
            tmem_ld_wait();
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float a = __uint_as_float(ov[i]) * inv, bb = __uint_as_float(ov[i + 1]) * inv;
                if (p.out_bf16) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(a, bb);
                    w[i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
                } else {
                    const __half2 h = __floats2half2_rn(a, bb);
                    w[i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
                }
            }
            uint4 *dst = reinterpret_cast<uint4 *>(static_cast<uint16_t *>(p.o) + orow + c * 32);
#pragma unroll
            for (int u = 0; u < 4; u++) dst[u] = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
        }
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<256>(tmem_base);
    }
}


}  // namespace
}  // namespace nb200

// q / k / v: fp16 [batch, heads, tokens, 128] contiguous, row-major inside a head (the PackQKV epilogue's layout); o: fp16 / bf16
// [batch, tokens_q, heads * 128].  tokens_q, tokens_kv multiples of 128.  scale: the softmax scale (1 / sqrt(128) for FLUX); the kernel works in
// base 2 like the reference (attention.cu:48-49).
extern "C" __attribute__((visibility("default"))) int nb200_attention_fp16(const void *q, const void *k, const void *v, void *o, int out_dtype, int batch,
                                                                            int heads, int tokens_q, int tokens_kv, float scale, void *stream_) {
    using namespace nb200;
    reset_launch_count();
    NB200_REQUIRE(q && k && v && o, "NULL tensor");
    NB200_REQUIRE(batch > 0 && heads > 0, "batch / heads must be positive");
    NB200_REQUIRE(tokens_q > 0 && tokens_q % kBM == 0, "tokens_q must be a positive multiple of 128 (attention.cu:51)");
    NB200_REQUIRE(tokens_kv > 0 && tokens_kv % kBN == 0, "tokens_kv must be a positive multiple of 128");
    NB200_REQUIRE(out_dtype == NB200_FP16 || out_dtype == NB200_BF16, "o must be fp16 or bf16");
    NB200_REQUIRE(static_cast<long long>(batch) * heads * tokens_q < (1ll << 31) && static_cast<long long>(batch) * heads * tokens_kv < (1ll << 31), "too many rows");
    if (int rc = nb200_check_device()) return rc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CUtensorMap tm_q, tm_k, tm_v;
    const uint64_t rows_q = static_cast<uint64_t>(batch) * heads * tokens_q, rows_kv = static_cast<uint64_t>(batch) * heads * tokens_kv;
    int rc = make_map_2d(&tm_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, q, kD, rows_q, kD * 2, 64, kBM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_2d(&tm_k, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, k, kD, rows_kv, kD * 2, 64, kBN, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_2d(&tm_v, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, v, kD, rows_kv, kD * 2, 64, kBN, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    AttnParams p;
    p.o = o;
    p.heads = heads;
    p.tokens_q = tokens_q;
    p.tokens_kv = tokens_kv;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.out_bf16 = out_dtype == NB200_BF16;
    const char *dbg_env = getenv("NB200_ATTN_DEBUG");
    p.debug = dbg_env ? atoi(dbg_env) : 0;
    const char *ver_env = getenv("NB200_ATTN_V");   // 1 = the first kernel (one CTA per SM, P through shared memory); read per launch (tools/attn_bench.py)
    if (ver_env && atoi(ver_env) == 1) {
        const size_t smem = sizeof(SmemA) + 1024;
        if (int rc2 = set_max_smem_once(reinterpret_cast<const void *>(attention_fp16_kernel), smem)) return rc2;
        LaunchCfg lc(dim3(tokens_q / kBM, heads, batch), dim3(kThreadsAttn), smem, stream);
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, attention_fp16_kernel, tm_q, tm_k, tm_v, p));
    } else {
        const size_t smem = sizeof(SmemA2) + 1024;
        if (int rc2 = set_max_smem_once(reinterpret_cast<const void *>(attention_fp16_v2_kernel), smem)) return rc2;
        LaunchCfg lc(dim3(tokens_q / kBM, heads, batch), dim3(kThreadsV2), smem, stream);
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, attention_fp16_v2_kernel, tm_q, tm_k, tm_v, p));
    }
    count_launch();
    return NB200_OK;
}
