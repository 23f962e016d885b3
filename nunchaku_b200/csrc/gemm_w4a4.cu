// Fused SVDQuant W4A4 GEMM for B200 (sm_100a): tcgen05 tensor cores, TMEM accumulators, TMA.
//
// Replaces gemm_w4a4_kernel / gemm_w4a4_fp4_kernel and their epilogue chain (reference
// src/kernels/zgemm/gemm_w4a4.cuh:358-405,1046-1095; lora.cuh:110-241; gemm_base.cuh:667-781;
// host dispatch gemm_w4a4_launch_impl.cuh:7-424).  See DESIGN.md section 4.2.
//
//   out[m, n] = act( (sum_k A[m,k] W[n,k]  +  sum_r La[m,r] Lu'[n,r]) * cscale[n] + bias[n] )
//
//   * NVFP4: A, W are e2m1 with ue4m3 scales per 16 -> tcgen05.mma kind::mxf4nvf4.block_scale
//     straight from TMA-staged packed tiles; scale factors go smem -> TMEM with tcgen05.cp.
//   * INT4 : there is no 4-bit integer kind on tcgen05.  Converter warps expand the
//     TMA-staged packed nibbles to the model's 16-bit float type with the per-group scales folded
//     in (exact integers times an hT scale, one rounding) into 128B-swizzled K-major tiles, and
//     the main loop is a single fp32 accumulation with kind::f16 -- no per-group TMEM drain.
//   * low-rank branch: La (fp32, converted to hT with lora_scales by the epilogue warps) times Lu'
//     (lora_up pre-divided by cscale at load) is accumulated into the SAME TMEM tile by a few
//     kind::f16 MMAs with K = rank, so the epilogue reads one accumulator and writes HBM once.
//   * persistent, warp specialised: warp0 TMA producer, warp1 MMA issuer, warp2 TMEM allocator,
//     warps4-7 epilogue (TMEM -> regs -> swizzled smem -> TMA store), INT4 adds 8 converter warps.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace nb200 {
namespace {

using namespace ptx;

constexpr int BM = 128;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiThreads = 128;
constexpr int kConvWarp0 = 8;
constexpr int kNumConvThreads = 256;
constexpr int kLoraChunk = 32;  // ranks per low-rank MMA group (2 x K16)

struct GemmParams {
    const uint8_t *sfa;     // FP4: activation scale tiles
    const uint8_t *sfb;     // FP4: weight scale tiles
    const void *ascales;    // INT4: hT [K/64][Mp]
    const void *wscales;    // INT4: hT [K/64][N]
    const float *bias;      // [N] or null
    const float *cscale;    // [N] or null
    const float *lora_act;  // [Mp][R] or null
    const void *lora_up;    // blocks [Rp/32][N/8][4][8][8] or null
    int Mp, N, K;
    int R, Rp;
    int M_out, N_out;
    int num_n_blocks, num_tiles;
    int mid_act;
    int act_unsigned;
    float lora_scales[NB200_MAX_LORA_SCALES];
};

template <bool FP4, int BN>
struct Cfg {
    static constexpr int kBK = FP4 ? 256 : 64;                 // k elements per pipeline stage
    static constexpr int kStages = FP4 ? (BN == 256 ? 3 : 4) : 4;  // TMA ring depth
    static constexpr int kConvStages = 2;                      // INT4: converted-tile ring depth
    static constexpr int kNumAcc = FP4 ? (BN <= 128 ? 2 : 1) : 2;
    static constexpr int kABytes = FP4 ? BM * 128 : BM * 32;   // packed A tile per stage
    static constexpr int kBBytes = FP4 ? BN * 128 : BN * 32;
    static constexpr int kSfaCols = 16;                         // 4 K64 blocks x 4 columns
    static constexpr int kSfbCols = BN / 8;                     // 4 K64 blocks x BN/32 columns
    static constexpr int kTmemSfa = kNumAcc * BN;
    static constexpr int kTmemSfb = kTmemSfa + kSfaCols;
    static constexpr int kThreads = FP4 ? 256 : 512;
    static_assert(!FP4 || kTmemSfb + kSfbCols <= 512, "TMEM budget");
    static_assert(kNumAcc * BN <= 512, "TMEM budget");
};

template <bool FP4, int BN>
struct alignas(1024) Smem {
    using C = Cfg<FP4, BN>;
    // TMA-staged packed operands
    alignas(1024) uint8_t a[C::kStages][C::kABytes];
    alignas(1024) uint8_t b[C::kStages][C::kBBytes];
    // FP4: scale-factor tiles (tcgen05.cp layout).  INT4: per-group scales (hT)
    alignas(128) uint8_t sa[C::kStages][FP4 ? 4 * 512 : BM * 2];
    alignas(128) uint8_t sb[C::kStages][FP4 ? (BN / 128) * 4 * 512 : BN * 2];
    // INT4: converted hT tiles, 128B-swizzled K-major [rows][64]
    alignas(1024) uint8_t a_cv[FP4 ? 1 : C::kConvStages][FP4 ? 16 : BM * 128];
    alignas(1024) uint8_t b_cv[FP4 ? 1 : C::kConvStages][FP4 ? 16 : BN * 128];
    // low-rank operands, UMMA no-swizzle K-major core-matrix order, one 32-rank chunk
    alignas(1024) uint8_t lora_a[BM * kLoraChunk * 2];
    alignas(1024) uint8_t lora_b[BN * kLoraChunk * 2];
    // epilogue staging for TMA store: [128 rows][64 cols] hT, 128B swizzle, double buffered
    alignas(1024) uint8_t out_stage[2][BM * 128];
    float bias[BN];
    float cscale[BN];
    uint64_t full[C::kStages];
    uint64_t empty[C::kStages];
    uint64_t cfull[C::kConvStages];
    uint64_t cempty[C::kConvStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint64_t lora_b_full;
    uint64_t lora_a_full;
    uint64_t lora_empty;
    uint32_t tmem_base;
};

struct PipeState {
    uint32_t idx = 0, phase = 0;
    __device__ __forceinline__ void advance(uint32_t n) {
        if (++idx == n) {
            idx = 0;
            phase ^= 1;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// INT4 converter: 16 packed bytes (32 elements) -> 64 bytes of hT in a 128B-swizzled row
// ---------------------------------------------------------------------------------------------
template <typename hT>
__device__ __forceinline__ void convert_unit(const uint8_t *pk_tile, uint8_t *cv_tile, int unit, const hT *scales,
                                             uint32_t offset_bits) {
    using Tr = HalfTraits<hT>;
    using T2 = typename Tr::T2;
    constexpr uint32_t kMagic = Tr::kIsBf16 ? 0x43004300u : 0x64006400u;  // 128 + u  |  1024 + u
    const int r = unit >> 1, h = unit & 1;
    const uint4 pk = *reinterpret_cast<const uint4 *>(pk_tile + unit * 16);
    const hT s = scales[r];
    T2 s2;
    s2.x = s;
    s2.y = s;
    const T2 off = *reinterpret_cast<const T2 *>(&offset_bits);
    const uint32_t words[4] = {pk.x, pk.y, pk.z, pk.w};
    uint8_t *row = cv_tile + r * 128;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        uint32_t o[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            uint32_t bits = ((words[w] >> (4 * p)) & 0x000F000Fu) | kMagic;  // elements 2p, 2p+1
            T2 v = *reinterpret_cast<T2 *>(&bits);
            v = __hmul2(__hsub2(v, off), s2);  // exact integer, then one rounding
            o[p] = *reinterpret_cast<uint32_t *>(&v);
        }
        const int chunk = (4 * h + w) ^ (r & 7);
        *reinterpret_cast<uint4 *>(row + chunk * 16) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool FP4, typename hT, int BN>
__global__ void __launch_bounds__(Cfg<FP4, BN>::kThreads, 1)
gemm_w4a4_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_wgt,
                 const __grid_constant__ CUtensorMap tm_out, const GemmParams p) {
    using C = Cfg<FP4, BN>;
    using S = Smem<FP4, BN>;
    using Tr = HalfTraits<hT>;
    extern __shared__ uint8_t smem_raw[];
    S &s = *reinterpret_cast<S *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int k64_total = p.K >> 6;
    const int num_kblocks = FP4 ? (k64_total + 3) >> 2 : k64_total;
    const int lora_chunks = p.lora_up != nullptr ? p.Rp / kLoraChunk : 0;

    // ---- one-time setup -----------------------------------------------------------------------
    if (warp == 0 && elect_one()) {
        prefetch_tensormap(&tm_act);
        prefetch_tensormap(&tm_wgt);
        prefetch_tensormap(&tm_out);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < C::kStages; i++) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], FP4 ? 1 : kNumConvThreads);
        }
        for (int i = 0; i < C::kConvStages; i++) {
            mbar_init(&s.cfull[i], kNumConvThreads);
            mbar_init(&s.cempty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&s.tmem_full[i], 1);
            mbar_init(&s.tmem_empty[i], kNumEpiThreads);
        }
        mbar_init(&s.lora_b_full, 1);
        mbar_init(&s.lora_a_full, kNumEpiThreads);
        mbar_init(&s.lora_empty, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(&s.tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s.tmem_base;

    if (warp == 0) {
        // =================================== TMA producer =======================================
        if (elect_one()) {
            PipeState st;
            uint32_t lora_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int mb = tile / p.num_n_blocks, nb = tile % p.num_n_blocks;
                const int m0 = mb * BM, n0 = nb * BN;
                for (int kb = 0; kb < num_kblocks; kb++) {
                    mbar_wait(&s.empty[st.idx], st.phase ^ 1);
                    if constexpr (FP4) {
                        const int nj = min(4, k64_total - 4 * kb);
                        const uint32_t sf_bytes = nj * 512;
                        mbar_expect_tx(&s.full[st.idx], C::kABytes + C::kBBytes + sf_bytes * (1 + BN / 128));
                        tma_load_2d(s.a[st.idx], &tm_act, &s.full[st.idx], kb * 128, m0);
                        tma_load_2d(s.b[st.idx], &tm_wgt, &s.full[st.idx], kb * 128, n0);
                        bulk_load(s.sa[st.idx], p.sfa + (static_cast<size_t>(mb) * k64_total + 4 * kb) * 512, sf_bytes,
                                  &s.full[st.idx]);
#pragma unroll
                        for (int h = 0; h < BN / 128; h++)
                            bulk_load(s.sb[st.idx] + h * 2048,
                                      p.sfb + (static_cast<size_t>(n0 / 128 + h) * k64_total + 4 * kb) * 512, sf_bytes,
                                      &s.full[st.idx]);
                    } else {
                        mbar_expect_tx(&s.full[st.idx], C::kABytes + C::kBBytes + BM * 2 + BN * 2);
                        tma_load_2d(s.a[st.idx], &tm_act, &s.full[st.idx], kb * 32, m0);
                        tma_load_2d(s.b[st.idx], &tm_wgt, &s.full[st.idx], kb * 32, n0);
                        bulk_load(s.sa[st.idx], reinterpret_cast<const hT *>(p.ascales) + static_cast<size_t>(kb) * p.Mp + m0,
                                  BM * 2, &s.full[st.idx]);
                        bulk_load(s.sb[st.idx], reinterpret_cast<const hT *>(p.wscales) + static_cast<size_t>(kb) * p.N + n0,
                                  BN * 2, &s.full[st.idx]);
                    }
                    st.advance(C::kStages);
                }
                for (int c = 0; c < lora_chunks; c++) {
                    mbar_wait(&s.lora_empty, lora_phase ^ 1);
                    mbar_expect_tx(&s.lora_b_full, BN * kLoraChunk * 2);
                    bulk_load(s.lora_b,
                              reinterpret_cast<const hT *>(p.lora_up) + (static_cast<size_t>(c) * p.N + n0) * kLoraChunk,
                              BN * kLoraChunk * 2, &s.lora_b_full);
                    lora_phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ==================================== MMA issuer ========================================
        if (elect_one()) {
            PipeState st;   // FP4: TMA ring.  INT4: converted ring
            uint32_t lora_phase = 0;
            uint32_t acc_phase[2] = {0, 0};
            int it = 0;
            constexpr uint32_t idesc_main = FP4 ? make_idesc_nvf4(BM, BN) : make_idesc_f16(Tr::kIsBf16, BM, BN);
            constexpr uint32_t idesc_lora = make_idesc_f16(Tr::kIsBf16, BM, BN);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++) {
                const int acc = it % C::kNumAcc;
                mbar_wait(&s.tmem_empty[acc], acc_phase[acc] ^ 1);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kblocks; kb++) {
                    if constexpr (FP4) {
                        mbar_wait(&s.full[st.idx], st.phase);
                        tc_fence_after_sync();
                        const int nj = min(4, k64_total - 4 * kb);
                        for (int j = 0; j < nj; j++) {
                            tc_cp_32x128b_warpx4(tmem_base + C::kTmemSfa + 4 * j,
                                                 make_smem_desc(smem_u32(s.sa[st.idx] + j * 512), 0, 128, kLayoutNoSwizzle));
#pragma unroll
                            for (int h = 0; h < BN / 128; h++)
                                tc_cp_32x128b_warpx4(
                                    tmem_base + C::kTmemSfb + (BN / 32) * j + 4 * h,
                                    make_smem_desc(smem_u32(s.sb[st.idx] + h * 2048 + j * 512), 0, 128, kLayoutNoSwizzle));
                        }
                        const uint32_t a_addr = smem_u32(s.a[st.idx]), b_addr = smem_u32(s.b[st.idx]);
                        for (int j = 0; j < nj; j++)
                            tc_mma_nvf4(tmem_d, make_sw128_kmajor_desc(a_addr + j * 32), make_sw128_kmajor_desc(b_addr + j * 32),
                                        idesc_main, tmem_base + C::kTmemSfa + 4 * j,
                                        tmem_base + C::kTmemSfb + (BN / 32) * j, (kb | j) != 0);
                        tc_commit(&s.empty[st.idx]);
                        st.advance(C::kStages);
                    } else {
                        mbar_wait(&s.cfull[st.idx], st.phase);
                        tc_fence_after_sync();
                        const uint32_t a_addr = smem_u32(s.a_cv[st.idx]), b_addr = smem_u32(s.b_cv[st.idx]);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            tc_mma_f16(tmem_d, make_sw128_kmajor_desc(a_addr + j * 32), make_sw128_kmajor_desc(b_addr + j * 32),
                                       idesc_main, (kb | j) != 0);
                        tc_commit(&s.cempty[st.idx]);
                        st.advance(C::kConvStages);
                    }
                }
                for (int c = 0; c < lora_chunks; c++) {
                    mbar_wait(&s.lora_b_full, lora_phase);
                    mbar_wait(&s.lora_a_full, lora_phase);
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(s.lora_a), b_addr = smem_u32(s.lora_b);
#pragma unroll
                    for (int j = 0; j < kLoraChunk / 16; j++)
                        tc_mma_f16(tmem_d, make_smem_desc(a_addr + j * 256, 128, 512, kLayoutNoSwizzle),
                                   make_smem_desc(b_addr + j * 256, 128, 512, kLayoutNoSwizzle), idesc_lora, 1);
                    tc_commit(&s.lora_empty);
                    lora_phase ^= 1;
                }
                tc_commit(&s.tmem_full[acc]);
                acc_phase[acc] ^= 1;
            }
        }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
        // ===================================== epilogue ==========================================
        const int q = warp - kEpiWarp0;        // TMEM lane quadrant
        const int et = threadIdx.x - kEpiWarp0 * 32;
        const int row = q * 32 + lane;         // row inside the tile
        uint32_t lora_phase = 0;
        uint32_t acc_phase[2] = {0, 0};
        int it = 0;
        uint32_t store_count = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, it++) {
            const int mb = tile / p.num_n_blocks, nb = tile % p.num_n_blocks;
            const int m0 = mb * BM, n0 = nb * BN;
            const int acc = it % C::kNumAcc;

            // per-tile channel vectors
            named_bar_sync(1, kNumEpiThreads);
            for (int i = et; i < BN; i += kNumEpiThreads) {
                s.bias[i] = p.bias != nullptr ? p.bias[n0 + i] : 0.f;
                s.cscale[i] = p.cscale != nullptr ? p.cscale[n0 + i] : 1.f;
            }
            named_bar_sync(1, kNumEpiThreads);

            // low-rank activations: fp32 -> * lora_scale -> hT  (lora.cuh:145-151), one row per thread
            for (int c = 0; c < lora_chunks; c++) {
                mbar_wait(&s.lora_empty, lora_phase ^ 1);
                const float *src = p.lora_act + static_cast<size_t>(m0 + row) * p.R + c * kLoraChunk;
                uint8_t *dst = s.lora_a + (row >> 3) * 512 + (row & 7) * 16;
#pragma unroll
                for (int o = 0; o < 4; o++) {  // rank octet
                    const int r0 = c * kLoraChunk + o * 8;
                    uint32_t w[4] = {0, 0, 0, 0};
                    if (r0 < p.R) {
                        const float4 f0 = *reinterpret_cast<const float4 *>(src + o * 8);
                        const float4 f1 = *reinterpret_cast<const float4 *>(src + o * 8 + 4);
                        const float sc = p.lora_scales[r0 >> 4];
                        typename Tr::T2 h0 = Tr::from_float2(make_float2(f0.x * sc, f0.y * sc));
                        typename Tr::T2 h1 = Tr::from_float2(make_float2(f0.z * sc, f0.w * sc));
                        typename Tr::T2 h2 = Tr::from_float2(make_float2(f1.x * sc, f1.y * sc));
                        typename Tr::T2 h3 = Tr::from_float2(make_float2(f1.z * sc, f1.w * sc));
                        w[0] = *reinterpret_cast<uint32_t *>(&h0);
                        w[1] = *reinterpret_cast<uint32_t *>(&h1);
                        w[2] = *reinterpret_cast<uint32_t *>(&h2);
                        w[3] = *reinterpret_cast<uint32_t *>(&h3);
                    }
                    *reinterpret_cast<uint4 *>(dst + o * 128) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                fence_proxy_async_smem();
                mbar_arrive(&s.lora_a_full);
                lora_phase ^= 1;
            }

            mbar_wait(&s.tmem_full[acc], acc_phase[acc]);
            acc_phase[acc] ^= 1;
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int ch = 0; ch < BN / 64; ch++) {
                uint32_t v0[32], v1[32];
                tmem_ld_32x32b_x32(taddr + ch * 64, v0);
                tmem_ld_32x32b_x32(taddr + ch * 64 + 32, v1);
                tmem_ld_wait();
                if (ch == BN / 64 - 1) {
                    tc_fence_before_sync();
                    mbar_arrive(&s.tmem_empty[acc]);
                }
                const int buf = store_count & 1;
                if (et == 0) bulk_wait_group_read<1>();
                named_bar_sync(2, kNumEpiThreads);
                uint8_t *srow = s.out_stage[buf] + row * 128;
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int col = c8 * 8 + 2 * i;  // within the 64-wide chunk
                        float y0 = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]);
                        float y1 = __uint_as_float(col + 1 < 32 ? v0[col + 1] : v1[col + 1 - 32]);
                        y0 = fmaf(y0, s.cscale[ch * 64 + col], s.bias[ch * 64 + col]);
                        y1 = fmaf(y1, s.cscale[ch * 64 + col + 1], s.bias[ch * 64 + col + 1]);
                        if (p.mid_act != NB200_ACT_NONE) {
                            // the reference applies the activation to the hT-rounded value
                            const float2 r = Tr::to_float2(Tr::from_float2(make_float2(y0, y1)));
                            if (p.mid_act == NB200_ACT_GELU) {
                                y0 = gelu_f32(r.x);
                                y1 = gelu_f32(r.y);
                            } else {
                                y0 = silu_f32(r.x);
                                y1 = silu_f32(r.y);
                            }
                        }
                        if constexpr (!Tr::kIsBf16) {  // fp16 stores clamp (gemm_base.cuh:688-696)
                            y0 = fminf(fmaxf(y0, -65504.f), 65504.f);
                            y1 = fminf(fmaxf(y1, -65504.f), 65504.f);
                        }
                        typename Tr::T2 h = Tr::from_float2(make_float2(y0, y1));
                        w[i] = *reinterpret_cast<uint32_t *>(&h);
                    }
                    *reinterpret_cast<uint4 *>(srow + ((c8 ^ (row & 7)) * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                fence_proxy_async_smem();
                named_bar_sync(3, kNumEpiThreads);
                if (et == 0) {
                    tma_store_2d(&tm_out, s.out_stage[buf], n0 + ch * 64, m0);
                    bulk_commit_group();
                }
                store_count++;
            }
        }
        if (et == 0) bulk_wait_group<0>();
    } else if (!FP4 && warp >= kConvWarp0) {
        // ============================ INT4 -> hT converter warps ===================================
        if constexpr (!FP4) {
            const int ct = threadIdx.x - kConvWarp0 * 32;
            PipeState pst, cst;
            // offset removed from the magic-biased value: 128/1024 (+8 when the nibble is offset-binary)
            typename Tr::T2 offA2, offB2;
            {
                const float base = Tr::kIsBf16 ? 128.f : 1024.f;
                const hT oa = Tr::from_float(base + (p.act_unsigned ? 0.f : 8.f));
                const hT ob = Tr::from_float(base + 8.f);
                offA2.x = oa;
                offA2.y = oa;
                offB2.x = ob;
                offB2.y = ob;
            }
            const uint32_t offA = *reinterpret_cast<uint32_t *>(&offA2);
            const uint32_t offB = *reinterpret_cast<uint32_t *>(&offB2);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < num_kblocks; kb++) {
                    mbar_wait(&s.full[pst.idx], pst.phase);
                    mbar_wait(&s.cempty[cst.idx], cst.phase ^ 1);
                    convert_unit<hT>(s.a[pst.idx], s.a_cv[cst.idx], ct, reinterpret_cast<const hT *>(s.sa[pst.idx]), offA);
#pragma unroll
                    for (int i = 0; i < BN / 128; i++)
                        convert_unit<hT>(s.b[pst.idx], s.b_cv[cst.idx], ct + i * kNumConvThreads,
                                         reinterpret_cast<const hT *>(s.sb[pst.idx]), offB);
                    fence_proxy_async_smem();
                    mbar_arrive(&s.cfull[cst.idx]);
                    mbar_arrive(&s.empty[pst.idx]);
                    pst.advance(C::kStages);
                    cst.advance(C::kConvStages);
                }
            }
        }
    }

    // ---- teardown -------------------------------------------------------------------------------
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
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

template <bool FP4, typename hT, int BN>
int launch(const nb200_gemm_args &a, cudaStream_t stream) {
    using C = Cfg<FP4, BN>;
    using S = Smem<FP4, BN>;
    CUtensorMap tm_act, tm_wgt, tm_out;
    const CUtensorMapSwizzle in_swz = FP4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const uint32_t in_box = FP4 ? 128 : 32;
    int rc = make_map_2d(&tm_act, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.act, a.K / 2, a.Mp, a.K / 2, in_box, BM, in_swz);
    if (rc) return rc;
    rc = make_map_2d(&tm_wgt, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.wgt, a.K / 2, a.N, a.K / 2, in_box, BN, in_swz);
    if (rc) return rc;
    const CUtensorMapDataType odt =
        HalfTraits<hT>::kIsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    rc = make_map_2d(&tm_out, odt, a.out, a.N_out, a.M_out, static_cast<uint64_t>(a.N_out) * 2, 64, BM,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;

    GemmParams p;
    p.sfa = static_cast<const uint8_t *>(a.ascales);
    p.sfb = static_cast<const uint8_t *>(a.wscales);
    p.ascales = a.ascales;
    p.wscales = a.wscales;
    p.bias = a.bias;
    p.cscale = a.cscale;
    p.lora_act = a.lora_act_in;
    p.lora_up = a.R_up > 0 ? a.lora_up : nullptr;
    p.Mp = a.Mp;
    p.N = a.N;
    p.K = a.K;
    p.R = a.R_up;
    p.Rp = (a.R_up + 31) / 32 * 32;
    p.M_out = a.M_out;
    p.N_out = a.N_out;
    p.num_n_blocks = a.N / BN;
    p.num_tiles = (a.Mp / BM) * p.num_n_blocks;
    p.mid_act = a.mid_act;
    p.act_unsigned = a.act_unsigned;
    for (int i = 0; i < NB200_MAX_LORA_SCALES; i++) p.lora_scales[i] = a.lora_scales[i];

    static int num_sms_cached = 0;
    if (num_sms_cached == 0) {
        int dev = 0;
        NB200_CUDA_CHECK(cudaGetDevice(&dev));
        NB200_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms_cached, cudaDevAttrMultiProcessorCount, dev));
    }
    const int num_sms = a.num_sms > 0 ? a.num_sms : num_sms_cached;
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    const size_t smem_bytes = sizeof(S) + 1024;
    auto kern = gemm_w4a4_kernel<FP4, hT, BN>;
    static bool attr_set = false;
    if (!attr_set) {
        NB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
        attr_set = true;
    }
    kern<<<grid, C::kThreads, smem_bytes, stream>>>(tm_act, tm_wgt, tm_out, p);
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    return NB200_OK;
}

template <bool FP4, typename hT>
int launch_bn(const nb200_gemm_args &a, cudaStream_t stream) {
    int bn = a.block_n;
    if (bn == 0) bn = (a.N % 256 == 0 && (a.Mp / BM) * (a.N / 256) >= 96) ? 256 : 128;
    if (bn == 256 && a.N % 256 == 0) return launch<FP4, hT, 256>(a, stream);
    if (bn == 128) return launch<FP4, hT, 128>(a, stream);
    return fail(NB200_ERR_INVALID_ARGUMENT, "block_n must be 0, 128 or 256 and divide N");
}

}  // namespace
}  // namespace nb200

extern "C" __attribute__((visibility("default"))) int nb200_gemm_w4a4(const nb200_gemm_args *a, void *stream_) {
    using namespace nb200;
    reset_launch_count();
    NB200_REQUIRE(a != nullptr, "args is NULL");
    NB200_REQUIRE(a->act && a->wgt && a->ascales && a->wscales, "act/wgt/ascales/wscales must be non-NULL");
    NB200_REQUIRE(a->Mp > 0 && a->Mp % 256 == 0, "Mp must be a positive multiple of 256");
    NB200_REQUIRE(a->N > 0 && a->N % 128 == 0, "N must be a positive multiple of 128");
    NB200_REQUIRE(a->K > 0 && a->K % 128 == 0, "K must be a positive multiple of 128");
    NB200_REQUIRE(a->dtype == NB200_FP16 || a->dtype == NB200_BF16, "dtype must be fp16 or bf16");
    NB200_REQUIRE(a->R_up >= 0 && a->R_up % 16 == 0, "R_up must be a multiple of 16");
    NB200_REQUIRE((a->R_up == 0) || (a->lora_act_in && a->lora_up), "lora_act_in and lora_up go together");
    NB200_REQUIRE(a->R_up <= 16 * NB200_MAX_LORA_SCALES, "rank exceeds MAX_RANK (1024)");
    if (a->qout != nullptr || a->rotary_emb != nullptr || a->lora_down_next != nullptr)
        return fail(NB200_ERR_UNSUPPORTED, "fused next-layer quantize / RMSNorm+RoPE epilogues are not built yet");
    NB200_REQUIRE(a->out != nullptr, "out must be non-NULL");
    NB200_REQUIRE(a->M_out > 0 && a->M_out <= a->Mp && a->Mp - a->M_out < 256, "M_out must be in (Mp-256, Mp]");
    NB200_REQUIRE(a->N_out > 0 && a->N_out <= a->N && a->N - a->N_out < 128, "N_out must be in (N-128, N]");
    NB200_REQUIRE(a->N_out % 8 == 0, "out row pitch must be a multiple of 16 bytes (TMA store)");
    NB200_REQUIRE((reinterpret_cast<uintptr_t>(a->act) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->wgt) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
                  "act/wgt/out must be 16-byte aligned");
    NB200_REQUIRE(a->fp4 || !a->cscale, "INT4 path has no per-channel scale (alpha == 1, launch_impl:107)");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (a->fp4) {
        NB200_REQUIRE(!a->act_unsigned, "act_unsigned is INT4 only");
        return a->dtype == NB200_BF16 ? launch_bn<true, __nv_bfloat16>(*a, stream) : launch_bn<true, __half>(*a, stream);
    }
    return a->dtype == NB200_BF16 ? launch_bn<false, __nv_bfloat16>(*a, stream) : launch_bn<false, __half>(*a, stream);
}
