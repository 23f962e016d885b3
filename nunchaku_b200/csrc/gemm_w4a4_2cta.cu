// CTA-pair (cta_group::2) version of the fused SVDQuant W4A4 GEMM: the default-epilogue work horse.
//
// Why pairs (profiles/r01x_gemm_*_ncu.txt): at full tensor rate every dense
// tcgen05 kind reads 96 B/clk of operands from shared memory for a 128x256 tile and TMA writes
// the same 96 B/clk -- more than the 128 B/clk one SM's shared memory delivers -- and a 222 KB
// smem budget cannot hide L2 latency at that rate.  A pair of SMs computing a 256x256 tile splits
// the B operand: each SM stages (and, for INT4, converts) only 128 of the 256 weight rows, the
// tensor cores of both SMs read both halves.  Per SM: half the B bytes, half the conversion work,
// one more pipeline stage.
//
//   cluster = 2 CTAs along M.  CTA r owns A rows [m0 + 128 r, +128) and B rows [n0 + 128 r, +128).
//   Leader (rank 0) warp 1 issues tcgen05.cp / tcgen05.mma .cta_group::2 and multicast commits.
//   Every "data is ready" barrier the leader waits on lives in the leader's shared memory and is
//   signalled by both CTAs (TMA cta_group::2 complete_tx, remote mbarrier.arrive); every "slot is
//   free" barrier exists in both CTAs and is signalled by a multicast tcgen05.commit.
//
// Same operand layouts, arithmetic and epilogue as gemm_w4a4.cu (EPI_DEFAULT only; the fused
// quantise / RoPE epilogues stay on the single-CTA kernel for now).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace nb200 {
namespace {

using namespace ptx;

constexpr int BM = 128;   // rows per CTA (pair: 256)
constexpr int BN = 256;   // columns per pair tile
constexpr int BH = 128;   // B rows staged per CTA
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiThreads = 128;
constexpr int kConvWarp0 = 8;
constexpr int kNumConvThreads = 256;
constexpr int kLoraChunk = 32;

struct Gemm2Params {
    const void *ascales;   // INT4: hT [K/64][Mp]
    const void *wscales;   // INT4: hT [K/64][N]
    const float *bias;
    const float *cscale;
    const float *lora_act;
    int has_lora;
    int Mp, N, K, R, Rp;
    int num_n_blocks, num_tiles;   // pair tiles
    int mid_act, act_unsigned;
    long long *prof;  // optional [grid][16] cycle counters (tools/gemm_prof.py)
    int debug;        // NB200_GEMM_DEBUG ablation bits (see gemm_w4a4.cu): 4 = no main-loop MMAs, 8 = no epilogue math/stores
    float lora_scales[NB200_MAX_LORA_SCALES];
};

template <bool FP4>
struct Cfg2 {
    static constexpr int kStages = FP4 ? 4 : 6;
    static constexpr int kConvStages = 3;
    static constexpr int kNumAcc = FP4 ? 1 : 2;
    static constexpr int kABytes = FP4 ? BM * 128 : BM * 32;
    static constexpr int kBBytes = FP4 ? BH * 128 : BH * 32;
    static constexpr int kSaBytes = FP4 ? 4 * 512 : BM * 2;        // FP4: SFA tiles of own rows
    static constexpr int kSbBytes = FP4 ? 2 * 4 * 512 : BH * 2;    // FP4: SFB of ALL 256 columns
    static constexpr int kTmemSfa = kNumAcc * BN;
    static constexpr int kTmemSfb = kTmemSfa + 16;
    static constexpr int kEpiGroups = FP4 ? 2 : 1;   // see gemm_w4a4.cu: FP4 tiles are epilogue bound
    static constexpr int kEpiThreads = 128 * kEpiGroups;
    static constexpr int kThreads = FP4 ? (4 + 4 * kEpiGroups) * 32 : 512;
    static constexpr int kSfSet = 16 + 32;   // SFA + SFB columns of one k-stage; two sets, see gemm_w4a4.cu
    static_assert(!FP4 || kTmemSfa + 2 * kSfSet <= 512, "TMEM budget");
};

template <bool FP4>
struct alignas(1024) Smem2 {
    using C = Cfg2<FP4>;
    alignas(1024) uint8_t a[C::kStages][C::kABytes];
    alignas(1024) uint8_t b[C::kStages][C::kBBytes];
    alignas(128) uint8_t sa[C::kStages][C::kSaBytes];
    alignas(128) uint8_t sb[C::kStages][C::kSbBytes];
    alignas(1024) uint8_t a_cv[FP4 ? 1 : C::kConvStages][FP4 ? 16 : BM * 128];
    alignas(1024) uint8_t b_cv[FP4 ? 1 : C::kConvStages][FP4 ? 16 : BH * 128];
    alignas(1024) uint8_t lora_a[BM * kLoraChunk * 2];
    alignas(1024) uint8_t lora_b[BH * kLoraChunk * 2];
    alignas(1024) uint8_t out_stage[2][BM * 128];
    float bias[BN];
    float cscale[BN];
    uint64_t full[C::kStages];       // FP4: leader only (both CTAs' TMA bytes).  INT4: local
    uint64_t empty[C::kStages];      // FP4: multicast commit.                    INT4: local converters
    uint64_t cfull[C::kConvStages];  // INT4: leader, 2 x 256 converter arrivals
    uint64_t cempty[C::kConvStages]; // INT4: multicast commit
    uint64_t tmem_full[2];           // multicast commit
    uint64_t tmem_empty[2];          // leader, 2 x 128 epilogue arrivals
    uint64_t lora_b_full;            // leader, TMA bytes of both halves
    uint64_t lora_a_full;            // leader, 2 x 128 epilogue arrivals
    uint64_t lora_empty;             // multicast commit
    uint32_t tmem_base;
};

// cycle accounting of the barrier waits (written out only when p.prof != nullptr)
#define NB200_TIMED(acc, stmt)            \
    do {                                  \
        const long long _t0 = clock64();  \
        stmt;                             \
        (acc) += clock64() - _t0;         \
    } while (0)

struct PipeState {
    uint32_t idx = 0, phase = 0;
    __device__ __forceinline__ void advance(uint32_t n) {
        if (++idx == n) {
            idx = 0;
            phase ^= 1;
        }
    }
};

template <typename hT>
__device__ __forceinline__ void convert_unit(const uint8_t *pk_tile, uint8_t *cv_tile, int unit, const hT *scales,
                                             uint32_t offset_bits) {
    using Tr = HalfTraits<hT>;
    using T2 = typename Tr::T2;
    constexpr uint32_t kMagic = Tr::kIsBf16 ? 0x43004300u : 0x64006400u;
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
            uint32_t bits = ((words[w] >> (4 * p)) & 0x000F000Fu) | kMagic;
            T2 v = *reinterpret_cast<T2 *>(&bits);
            v = __hmul2(__hsub2(v, off), s2);
            o[p] = *reinterpret_cast<uint32_t *>(&v);
        }
        const int chunk = (4 * h + w) ^ (r & 7);
        *reinterpret_cast<uint4 *>(row + chunk * 16) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// wait on a barrier whose arrivals come from the other CTA as well (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("nb200: cluster mbarrier watchdog block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x,
                   smem_u32(bar), parity);
            __trap();
        }
    }
}

template <bool FP4, typename hT>
__global__ void __launch_bounds__(Cfg2<FP4>::kThreads, 1)
gemm_w4a4_2cta_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_wgt,
                      const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_sfa,
                      const __grid_constant__ CUtensorMap tm_sfb, const __grid_constant__ CUtensorMap tm_lu,
                      const Gemm2Params p) {
    using C = Cfg2<FP4>;
    using S = Smem2<FP4>;
    using Tr = HalfTraits<hT>;
    extern __shared__ uint8_t smem_raw[];
    // align inside the SHARED address space (pointer arithmetic on the __shared__ array): a round trip through uintptr_t
    // makes every access to `s` a generic LD/ST instead of LDS/STS (seen in the ncu source view)
    S &s = *reinterpret_cast<S *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

    const long long t_kernel0 = clock64();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int k64_total = p.K >> 6;
    const int num_kblocks = FP4 ? (k64_total + 3) >> 2 : k64_total;
    const int lora_chunks = p.has_lora ? p.Rp / kLoraChunk : 0;

    if (warp == 0 && elect_one()) {
        prefetch_tensormap(&tm_act);
        prefetch_tensormap(&tm_wgt);
        prefetch_tensormap(&tm_out);
        if (p.has_lora) prefetch_tensormap(&tm_lu);
        if constexpr (FP4) {
            prefetch_tensormap(&tm_sfa);
            prefetch_tensormap(&tm_sfb);
        }
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < C::kStages; i++) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], FP4 ? 1 : kNumConvThreads / 32);
        }
        for (int i = 0; i < C::kConvStages; i++) {
            mbar_init(&s.cfull[i], 2 * kNumConvThreads / 32);
            mbar_init(&s.cempty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&s.tmem_full[i], 1);
            mbar_init(&s.tmem_empty[i], 2 * C::kEpiThreads);
        }
        mbar_init(&s.lora_b_full, 1);
        mbar_init(&s.lora_a_full, 2 * C::kEpiThreads);
        mbar_init(&s.lora_empty, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_cg2<512>(&s.tmem_base);
    tc_fence_before_sync();
    cluster_sync();
    tc_fence_after_sync();
    const uint32_t tmem_base = s.tmem_base;
    const long long t_setup = clock64() - t_kernel0;
    griddep_launch_dependents();
    griddep_wait();

    if (warp == 0) {
        // =================================== TMA producer (both CTAs) ===============================
        if (elect_one()) {
            PipeState st;
            uint32_t lora_phase = 0;
            long long t_empty = 0;
            for (int tile = pair; tile < p.num_tiles; tile += num_pairs) {
                const int mb2 = tile / p.num_n_blocks, nb = tile % p.num_n_blocks;
                const int m0 = mb2 * 2 * BM + rank * BM;   // this CTA's A rows
                const int n0 = nb * BN;
                const int nh = n0 + rank * BH;             // this CTA's B rows
                for (int kb = 0; kb < num_kblocks; kb++) {
                    NB200_TIMED(t_empty, mbar_wait(&s.empty[st.idx], st.phase ^ 1));
                    if constexpr (FP4) {
                        // the leader's barrier collects the bytes of both CTAs
                        // (scale-factor boxes are always 4 tiles = 2 KB, also on the K tail)
                        if (leader) mbar_expect_tx(&s.full[st.idx], 2 * (C::kABytes + C::kBBytes + 3 * 2048));
                        tma_load_2d_cg2(s.a[st.idx], &tm_act, &s.full[st.idx], kb * 128, m0);
                        tma_load_2d_cg2(s.b[st.idx], &tm_wgt, &s.full[st.idx], kb * 128, nh);
                        // scale-factor tiles: rows of 128 x u32 (512 B); SFB for BOTH column halves
                        tma_load_2d_cg2(s.sa[st.idx], &tm_sfa, &s.full[st.idx], 0, (m0 / 128) * k64_total + 4 * kb);
                        tma_load_2d_cg2(s.sb[st.idx], &tm_sfb, &s.full[st.idx], 0, (n0 / 128) * k64_total + 4 * kb);
                        tma_load_2d_cg2(s.sb[st.idx] + 2048, &tm_sfb, &s.full[st.idx], 0, (n0 / 128 + 1) * k64_total + 4 * kb);
                    } else {
                        mbar_expect_tx(&s.full[st.idx], C::kABytes + C::kBBytes + BM * 2 + BH * 2);
                        tma_load_2d(s.a[st.idx], &tm_act, &s.full[st.idx], kb * 32, m0);
                        tma_load_2d(s.b[st.idx], &tm_wgt, &s.full[st.idx], kb * 32, nh);
                        bulk_load(s.sa[st.idx], reinterpret_cast<const hT *>(p.ascales) + static_cast<size_t>(kb) * p.Mp + m0,
                                  BM * 2, &s.full[st.idx]);
                        bulk_load(s.sb[st.idx], reinterpret_cast<const hT *>(p.wscales) + static_cast<size_t>(kb) * p.N + nh,
                                  BH * 2, &s.full[st.idx]);
                    }
                    st.advance(C::kStages);
                }
                for (int c = 0; c < lora_chunks; c++) {
                    mbar_wait(&s.lora_empty, lora_phase ^ 1);
                    if (leader) mbar_expect_tx(&s.lora_b_full, 2 * BH * kLoraChunk * 2);
                    // lora_up blocks [Rp/32][N/8][4][8][8] viewed as rows of 256 hT: row = chunk * N/8 + n/8
                    tma_load_2d_cg2(s.lora_b, &tm_lu, &s.lora_b_full, 0, c * (p.N >> 3) + (nh >> 3));
                    lora_phase ^= 1;
                }
            }
            if (p.prof) p.prof[blockIdx.x * 16 + 0] = t_empty;
        }
    } else if (warp == 1) {
        // ==================================== MMA issuer (leader CTA) ================================
        if (leader && elect_one()) {
            PipeState st;
            uint32_t lora_phase = 0;
            uint32_t acc_phase[2] = {0, 0};
            int it = 0;
            long long t_tmem_empty = 0, t_full = 0, t_lora = 0, t_first = 0;
            const long long t_mma0 = clock64();
            constexpr uint32_t idesc_main = FP4 ? make_idesc_nvf4(2 * BM, BN) : make_idesc_f16(Tr::kIsBf16, 2 * BM, BN);
            constexpr uint32_t idesc_lora = make_idesc_f16(Tr::kIsBf16, 2 * BM, BN);
            for (int tile = pair; tile < p.num_tiles; tile += num_pairs, it++) {
                const int acc = it % C::kNumAcc;
                NB200_TIMED(t_tmem_empty, mbar_wait_cluster(&s.tmem_empty[acc], acc_phase[acc] ^ 1));
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kblocks; kb++) {
                    if constexpr (FP4) {
                        NB200_TIMED(t_full, mbar_wait(&s.full[st.idx], st.phase));   // TMA barrier: CTA-scope acquire (see gemm_nvfp4_cluster.cu)
                        if (t_first == 0) t_first = clock64() - t_mma0;
                        tc_fence_after_sync();
                        const int nj = min(4, k64_total - 4 * kb);
                        const uint32_t sf_set = tmem_base + (kb & 1) * C::kSfSet;
                        for (int j = 0; j < nj; j++) {
                            tc_cp_32x128b_warpx4_cg2(sf_set + C::kTmemSfa + 4 * j,
                                                     make_smem_desc(smem_u32(s.sa[st.idx] + j * 512), 0, 128, kLayoutNoSwizzle));
#pragma unroll
                            for (int h = 0; h < 2; h++)
                                tc_cp_32x128b_warpx4_cg2(
                                    sf_set + C::kTmemSfb + 8 * j + 4 * h,
                                    make_smem_desc(smem_u32(s.sb[st.idx] + h * 2048 + j * 512), 0, 128, kLayoutNoSwizzle));
                        }
                        const uint32_t a_addr = smem_u32(s.a[st.idx]), b_addr = smem_u32(s.b[st.idx]);
                        for (int j = 0; j < ((p.debug & 4) ? 0 : nj); j++)
                            tc_mma_nvf4_cg2(tmem_d, make_sw128_kmajor_desc(a_addr + j * 32), make_sw128_kmajor_desc(b_addr + j * 32),
                                            idesc_main, sf_set + C::kTmemSfa + 4 * j, sf_set + C::kTmemSfb + 8 * j,
                                            (kb | j) != 0);
                        tc_commit_cg2(&s.empty[st.idx], 3);
                        st.advance(C::kStages);
                    } else {
                        NB200_TIMED(t_full, mbar_wait_cluster(&s.cfull[st.idx], st.phase));
                        if (t_first == 0) t_first = clock64() - t_mma0;
                        tc_fence_after_sync();
                        const uint32_t a_addr = smem_u32(s.a_cv[st.idx]), b_addr = smem_u32(s.b_cv[st.idx]);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            tc_mma_f16_cg2(tmem_d, make_sw128_kmajor_desc(a_addr + j * 32), make_sw128_kmajor_desc(b_addr + j * 32),
                                           idesc_main, (kb | j) != 0);
                        tc_commit_cg2(&s.cempty[st.idx], 3);
                        st.advance(C::kConvStages);
                    }
                }
                for (int c = 0; c < lora_chunks; c++) {
                    NB200_TIMED(t_lora, mbar_wait_cluster(&s.lora_b_full, lora_phase));
                    NB200_TIMED(t_lora, mbar_wait_cluster(&s.lora_a_full, lora_phase));
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(s.lora_a), b_addr = smem_u32(s.lora_b);
#pragma unroll
                    for (int j = 0; j < kLoraChunk / 16; j++)
                        tc_mma_f16_cg2(tmem_d, make_smem_desc(a_addr + j * 256, 128, 512, kLayoutNoSwizzle),
                                       make_smem_desc(b_addr + j * 256, 128, 512, kLayoutNoSwizzle), idesc_lora, 1);
                    tc_commit_cg2(&s.lora_empty, 3);
                    lora_phase ^= 1;
                }
                tc_commit_cg2(&s.tmem_full[acc], 3);
                acc_phase[acc] ^= 1;
            }
            if (p.prof) {
                p.prof[blockIdx.x * 16 + 1] = t_tmem_empty;
                p.prof[blockIdx.x * 16 + 2] = t_full;
                p.prof[blockIdx.x * 16 + 3] = t_lora;
                p.prof[blockIdx.x * 16 + 10] = clock64() - t_mma0;
                p.prof[blockIdx.x * 16 + 13] = t_first;
            }
        }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4 * C::kEpiGroups) {
        // ===================================== epilogue (both CTAs) ==================================
        constexpr int H = C::kEpiGroups;
        constexpr int CH = BN / 64;
        const int h = (warp - kEpiWarp0) >> 2;
        const int q = warp & 3;
        const int et = threadIdx.x - (kEpiWarp0 + 4 * h) * 32;
        const int eta = threadIdx.x - kEpiWarp0 * 32;
        const int row = q * 32 + lane;
        uint32_t lora_phase = 0;
        uint32_t acc_phase[2] = {0, 0};
        int it = 0;
        uint32_t store_count = 0;
        long long t_tmem_full = 0, t_pre = 0;
        const long long t_epi0 = clock64();
        const uint32_t lora_a_full_leader = mapa(smem_u32(&s.lora_a_full), 0);
        for (int tile = pair; tile < p.num_tiles; tile += num_pairs, it++) {
            const int mb2 = tile / p.num_n_blocks, nb = tile % p.num_n_blocks;
            const int m0 = mb2 * 2 * BM + rank * BM, n0 = nb * BN;
            const int acc = it % C::kNumAcc;
            const long long t_tile0 = clock64();

            named_bar_sync(1, C::kEpiThreads);
            for (int i = eta; i < BN; i += C::kEpiThreads) {
                s.bias[i] = p.bias != nullptr ? p.bias[n0 + i] : 0.f;
                s.cscale[i] = p.cscale != nullptr ? p.cscale[n0 + i] : 1.f;
            }
            named_bar_sync(1, C::kEpiThreads);

            for (int c = 0; c < lora_chunks; c++) {
                mbar_wait(&s.lora_empty, lora_phase ^ 1);
                const float *src = p.lora_act + static_cast<size_t>(m0 + row) * p.R + c * kLoraChunk;
                uint8_t *dst = s.lora_a + (row >> 3) * 512 + (row & 7) * 16;
#pragma unroll
                for (int oo = 0; oo < 4 / H; oo++) {
                    const int o = h * (4 / H) + oo;
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
                mbar_arrive_remote(lora_a_full_leader);
                lora_phase ^= 1;
            }

            t_pre += clock64() - t_tile0;
            NB200_TIMED(t_tmem_full, mbar_wait(&s.tmem_full[acc], acc_phase[acc]));
            acc_phase[acc] ^= 1;
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
            const uint32_t tmem_empty_leader = mapa(smem_u32(&s.tmem_empty[acc]), 0);
            // one 64-column chunk: scale/bias/activation, hT pack into the swizzled staging tile, TMA store
            auto do_chunk = [&](const int ch, const uint32_t (&v0)[32], const uint32_t (&v1)[32]) {
                const int buf = H == 2 ? h : (store_count & 1);
                if (et == 0) {
                    if constexpr (H == 2)
                        bulk_wait_group_read<0>();
                    else
                        bulk_wait_group_read<1>();
                }
                named_bar_sync(2 + 2 * h, kNumEpiThreads);
                uint8_t *srow = s.out_stage[buf] + row * 128;
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int col = c8 * 8 + 2 * i;
                        float y0 = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]);
                        float y1 = __uint_as_float(col + 1 < 32 ? v0[col + 1] : v1[col + 1 - 32]);
                        y0 = fmaf(y0, s.cscale[ch * 64 + col], s.bias[ch * 64 + col]);
                        y1 = fmaf(y1, s.cscale[ch * 64 + col + 1], s.bias[ch * 64 + col + 1]);
                        if (p.mid_act != NB200_ACT_NONE) {
                            const float2 r = Tr::to_float2(Tr::from_float2(make_float2(y0, y1)));
                            if (p.mid_act == NB200_ACT_GELU) {
                                y0 = gelu_f32(r.x);
                                y1 = gelu_f32(r.y);
                            } else {
                                y0 = silu_f32(r.x);
                                y1 = silu_f32(r.y);
                            }
                        }
                        if constexpr (!Tr::kIsBf16) {
                            y0 = fminf(fmaxf(y0, -65504.f), 65504.f);
                            y1 = fminf(fmaxf(y1, -65504.f), 65504.f);
                        }
                        typename Tr::T2 h = Tr::from_float2(make_float2(y0, y1));
                        w[i] = *reinterpret_cast<uint32_t *>(&h);
                    }
                    *reinterpret_cast<uint4 *>(srow + ((c8 ^ (row & 7)) * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                fence_proxy_async_smem();
                named_bar_sync(3 + 2 * h, kNumEpiThreads);
                if (et == 0) {
                    tma_store_2d(&tm_out, s.out_stage[buf], n0 + ch * 64, m0);
                    bulk_commit_group();
                }
                store_count++;
            };
            // FP4 with a single 256-column accumulator: pull this group's two chunks into registers first and
            // hand the accumulator back to the MMA warp before doing any math (the exposed part of the epilogue
            // shrinks from "drain + math + stores" to four tcgen05.ld).
            constexpr bool kEarlyRelease = FP4 && C::kNumAcc == 1 && CH / H == 2;
            if constexpr (kEarlyRelease) {
                uint32_t va[32], vb[32], vc[32], vd[32];
                const int ch0 = h * 2;
                tmem_ld_32x32b_x32(taddr + ch0 * 64, va);
                tmem_ld_32x32b_x32(taddr + ch0 * 64 + 32, vb);
                tmem_ld_32x32b_x32(taddr + ch0 * 64 + 64, vc);
                tmem_ld_32x32b_x32(taddr + ch0 * 64 + 96, vd);
                tmem_ld_wait();
                tc_fence_before_sync();
                mbar_arrive_remote(tmem_empty_leader);
                if (!(p.debug & 8)) {
                    do_chunk(ch0, va, vb);
                    do_chunk(ch0 + 1, vc, vd);
                }
            } else {
#pragma unroll 1
                for (int cc = 0; cc < CH / H; cc++) {
                    const int ch = h * (CH / H) + cc;
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32b_x32(taddr + ch * 64, v0);
                    tmem_ld_32x32b_x32(taddr + ch * 64 + 32, v1);
                    tmem_ld_wait();
                    if (cc == CH / H - 1) {
                        tc_fence_before_sync();
                        mbar_arrive_remote(tmem_empty_leader);
                    }
                    if (!(p.debug & 8)) do_chunk(ch, v0, v1);
                }
            }
        }
        if (et == 0) bulk_wait_group<0>();
        if (p.prof && eta == 0) {
            p.prof[blockIdx.x * 16 + 4] = t_tmem_full;
            p.prof[blockIdx.x * 16 + 5] = clock64() - t_epi0;
            p.prof[blockIdx.x * 16 + 11] = t_pre;
        }
    } else if (!FP4 && warp >= kConvWarp0) {
        // ============================ INT4 -> hT converter warps (both CTAs) ===========================
        if constexpr (!FP4) {
            const int ct = threadIdx.x - kConvWarp0 * 32;
            PipeState pst, cst;
            long long t_cfull = 0, t_cempty = 0, t_conv = 0;
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
            for (int tile = pair; tile < p.num_tiles; tile += num_pairs) {
                for (int kb = 0; kb < num_kblocks; kb++) {
                    NB200_TIMED(t_cfull, mbar_wait(&s.full[pst.idx], pst.phase));
                    NB200_TIMED(t_cempty, mbar_wait(&s.cempty[cst.idx], cst.phase ^ 1));
                    const long long t_c0 = clock64();
                    convert_unit<hT>(s.a[pst.idx], s.a_cv[cst.idx], ct, reinterpret_cast<const hT *>(s.sa[pst.idx]), offA);
                    convert_unit<hT>(s.b[pst.idx], s.b_cv[cst.idx], ct, reinterpret_cast<const hT *>(s.sb[pst.idx]), offB);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {  // one arrival per warp: 512 per-thread (remote) arrivals per k-block serialise on one word
                        mbar_arrive_remote(mapa(smem_u32(&s.cfull[cst.idx]), 0));
                        mbar_arrive(&s.empty[pst.idx]);
                    }
                    t_conv += clock64() - t_c0;
                    pst.advance(C::kStages);
                    cst.advance(C::kConvStages);
                }
            }
            if (p.prof && ct == 0) {
                p.prof[blockIdx.x * 16 + 6] = t_cfull;
                p.prof[blockIdx.x * 16 + 7] = t_cempty;
                p.prof[blockIdx.x * 16 + 8] = t_conv;
            }
        }
    }

    // both CTAs must stay alive until the pair's MMAs and remote arrivals are done
    tc_fence_before_sync();
    cluster_sync();
    if (p.prof && threadIdx.x == 0) {
        p.prof[blockIdx.x * 16 + 9] = clock64() - t_kernel0;
        p.prof[blockIdx.x * 16 + 12] = t_setup;
    }
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc_cg2<512>(tmem_base);
    }
}

template <bool FP4, typename hT>
int launch2(const nb200_gemm_args &a, cudaStream_t stream) {
    using C = Cfg2<FP4>;
    using S = Smem2<FP4>;
    CUtensorMap tm_act, tm_wgt, tm_out, tm_sfa, tm_sfb, tm_lu;
    const CUtensorMapSwizzle in_swz = FP4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const uint32_t in_box = FP4 ? 128 : 32;
    int rc = make_map_2d(&tm_act, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.act, a.K / 2, a.Mp, a.K / 2, in_box, BM, in_swz);
    if (rc) return rc;
    rc = make_map_2d(&tm_wgt, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.wgt, a.K / 2, a.N, a.K / 2, in_box, BH, in_swz);
    if (rc) return rc;
    const CUtensorMapDataType odt = HalfTraits<hT>::kIsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    rc = make_map_2d(&tm_out, odt, a.out, a.N_out, a.M_out, static_cast<uint64_t>(a.N_out) * 2, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    tm_sfa = tm_act;
    tm_sfb = tm_act;
    tm_lu = tm_act;
    if (FP4) {
        const uint64_t k64 = a.K / 64;
        rc = make_map_2d(&tm_sfa, CU_TENSOR_MAP_DATA_TYPE_UINT32, a.ascales, 128, (a.Mp / 128) * k64, 512, 128, 4, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        rc = make_map_2d(&tm_sfb, CU_TENSOR_MAP_DATA_TYPE_UINT32, a.wscales, 128, (a.N / 128) * k64, 512, 128, 4, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    const int Rp = (a.R_up + 31) / 32 * 32;
    if (a.R_up > 0) {
        rc = make_map_2d(&tm_lu, odt, a.lora_up, 256, static_cast<uint64_t>(Rp / 32) * (a.N / 8), 512, 256, BH / 8, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    Gemm2Params p;
    p.ascales = a.ascales;
    p.wscales = a.wscales;
    p.bias = a.bias;
    p.cscale = a.cscale;
    p.lora_act = a.lora_act_in;
    p.has_lora = a.R_up > 0;
    p.Mp = a.Mp;
    p.N = a.N;
    p.K = a.K;
    p.R = a.R_up;
    p.Rp = Rp;
    p.num_n_blocks = a.N / BN;
    p.num_tiles = (a.Mp / (2 * BM)) * p.num_n_blocks;
    p.mid_act = a.mid_act;
    p.act_unsigned = a.act_unsigned;
    p.prof = static_cast<long long *>(a.prof);
    static const int dbg = getenv("NB200_GEMM_DEBUG") ? atoi(getenv("NB200_GEMM_DEBUG")) : 0;
    p.debug = dbg;
    for (int i = 0; i < NB200_MAX_LORA_SCALES; i++) p.lora_scales[i] = a.lora_scales[i];

    int num_sms_dev = 0;
    if (int rc2 = current_device_sms(&num_sms_dev)) return rc2;
    const int num_sms = a.num_sms > 0 ? a.num_sms : num_sms_dev;
    const int pairs = p.num_tiles < num_sms / 2 ? p.num_tiles : num_sms / 2;
    const size_t smem_bytes = sizeof(S) + 1024;
    auto kern = gemm_w4a4_2cta_kernel<FP4, hT>;
    if (int rc2 = set_max_smem_once(reinterpret_cast<const void *>(kern), smem_bytes)) return rc2;
    LaunchCfg lc(dim3(2 * pairs), dim3(C::kThreads), smem_bytes, stream, 2);
    NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, tm_act, tm_wgt, tm_out, tm_sfa, tm_sfb, tm_lu, p));
    count_launch();
    return NB200_OK;
}

}  // namespace

// EPI_DEFAULT GEMM on CTA pairs; requires N % 256 == 0 (Mp % 256 == 0 always holds)
int gemm_w4a4_2cta_dispatch(const nb200_gemm_args &a, cudaStream_t stream) {
    if (a.fp4) return a.dtype == NB200_BF16 ? launch2<true, __nv_bfloat16>(a, stream) : launch2<true, __half>(a, stream);
    return a.dtype == NB200_BF16 ? launch2<false, __nv_bfloat16>(a, stream) : launch2<false, __half>(a, stream);
}

}  // namespace nb200
