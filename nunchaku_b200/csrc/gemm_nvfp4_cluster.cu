// NVFP4 fused SVDQuant GEMM on thread-block clusters: the default NVFP4 path for N % 256 == 0.
//
// Why (profiles/r02_*): at the measured NVFP4 tensor rate (tools/ubench/mma_peak.cu: ~6.9 PFLOP/s
// sustained) the kernel is bound by the L2 -> SM operand feed (~50 B/clk/SM chip-wide), so the lever is bytes per FLOP:
//
//   single CTA 128x256 tile                    A 16K + B 32K + SF 6K  = 54 KB per 16.8 MFLOP   (0.0033 B/FLOP)
//   CTA pair (cta_group::2) 256x256 tile       A 16K + B 16K + SF 6K  = 38 KB per CTA          (0.0023 B/FLOP)
//   2 pairs in one cluster, A multicast        A  8K + B 16K + SF 3K  = 27 KB per CTA          (0.0016 B/FLOP)
//
//   cluster = kPairs CTA pairs along N.  CTA (pi, q): pair pi computes the 256x256 tile (mb2, nb = nbc * kPairs + pi);
//   rank-in-pair q owns A rows [256 mb2 + 128 q, +128) and B rows [256 nb + 128 q, +128).
//   * B half: unicast into the CTA's own shared memory.
//   * A rows are the same for both pairs: CTA (pi, q) loads 64 of the 128 rows and MULTICASTS them to (0, q) and (1, q);
//     the activation scale tiles likewise (2 of the 4 K64 blocks each).
//   * weight scale factors cover all 256 columns of the pair tile and are needed in both CTAs of a pair (each SM's tensor
//     core scales its own 128 rows x 256 columns): CTA q loads the tile of columns [128 q, +128) and multicasts to its pair.
//   * every byte of a pair's stage is credited to the pair leader's `full` mbarrier (cta_group::2 TMA); a stage slot is
//     free again when the MMAs of BOTH pairs that read it have retired: each leader's tcgen05.commit is multicast to the
//     `empty` barrier of every CTA in the cluster (count = kPairs).
//
// One 256-column fp32 accumulator per CTA (256 + 2 x 48 scale-factor columns of TMEM): the epilogue warps pull their
// 128 columns into registers, hand TMEM back (one mbarrier arrival per WARP -- 16 per pair) and only then do the math.
// Epilogues: bias / per-channel scale / low-rank up (same TMEM tile) / SiLU / GELU / store, and RMSNorm + RoPE
// (+ PackQKV) with one 128-wide head per epilogue group.  Arithmetic identical to gemm_w4a4.cu.
//
// Replaces (reference): gemm_w4a4_fp4_kernel + epilogue chain, src/kernels/zgemm/gemm_w4a4.cuh:273-405,
// lora.cuh:110-241, gemm_base.cuh:667-781, epilogues.cuh:269-550.
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace nb200 {

int rope_inplace_dispatch(int dtype, void *qkv, int M, int N, const void *norm_q, const void *norm_k, const float *rotary, cudaStream_t stream);
int rope_pack_dispatch(int dtype, const void *qkv, int Mp, int N, const void *norm_q, const void *norm_k, const float *rotary, void *out_q, void *out_k,
                       void *out_v, long long sq, long long sk, long long sv, int attn_tokens, cudaStream_t stream);
namespace {

using namespace ptx;

constexpr int BM = 128;   // rows per CTA (pair: 256)
constexpr int BN = 256;   // columns per pair tile
constexpr int BH = 128;   // B rows staged per CTA
constexpr int kStages = 4;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiGroups = 2;
constexpr int kEpiThreads = 128 * kEpiGroups;
constexpr int kEpiWarps = 4 * kEpiGroups;
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;
constexpr int kLoraChunk = 32;
constexpr int kABytes = BM * 128, kBBytes = BH * 128, kSaBytes = 4 * 512, kSbBytes = 2 * 4 * 512;
constexpr int kStageBytes = kABytes + kBBytes + kSaBytes + kSbBytes;
constexpr int kTmemSfa = BN, kSfSet = 16 + 32;
static_assert(kTmemSfa + 2 * kSfSet <= 512, "TMEM budget");

enum { EPI_DEFAULT = 0, EPI_ROPE = 2 };

struct ClusterParams {
    const float *bias;
    const float *cscale;
    const float *lora_act;
    int has_lora;
    int Mp, N, K, R, Rp;
    int nct_n, num_ct;           // cluster tiles (256 rows x 256 kPairs columns) per row / in total
    int mid_act;
    // EPI_ROPE
    const void *norm_q, *norm_k;
    const float *rotary;
    __half *out_qkv[3];
    long long stride_head[3];
    int attn_tokens;
    void *out;                   // hT [M_out, N_out] row-major (null with PackQKV)
    int M_out, N_out;
    long long *prof;
    int debug;
    float lora_scales[NB200_MAX_LORA_SCALES];
};

struct alignas(1024) SmemC {
    alignas(1024) uint8_t a[kStages][kABytes];
    alignas(1024) uint8_t b[kStages][kBBytes];
    alignas(128) uint8_t sa[kStages][kSaBytes];
    alignas(128) uint8_t sb[kStages][kSbBytes];
    alignas(1024) uint8_t lora_a[2][BM * kLoraChunk * 2];   // low-rank activations of tile it in buffer it & 1 (converted one tile ahead)
    alignas(1024) uint8_t lora_b[BH * kLoraChunk * 2];
    alignas(1024) uint8_t out_stage[2][BM * 128];
    alignas(16) float bias[2][BN];     // tile it in buffer it & 1: fetched one tile ahead, read back as float4 broadcasts
    alignas(16) float cscale[2][BN];
    float normw[256];          // EPI_ROPE: RMSNorm weights (q | k)
    uint64_t full[kStages];    // pair leader: all TMA bytes of the pair's stage
    uint64_t empty[kStages];   // every CTA: kPairs multicast commits
    uint64_t tmem_full;        // multicast commit to the pair
    uint64_t tmem_empty;       // pair leader: 2 x 8 epilogue-warp arrivals
    uint64_t lora_b_full;      // pair leader: TMA bytes of both halves
    uint64_t lora_a_full[2];   // pair leader: 2 x 8 epilogue-warp arrivals, one barrier per lora_a buffer
    uint64_t lora_empty;       // multicast commit to the pair
    uint32_t tmem_base;
};
static_assert(sizeof(SmemC) + 1024 <= 232448, "shared memory budget (227 KB per CTA)");

#define NB200_TIMED(acc, stmt)            \
    do {                                  \
        const long long _t0 = clock64();  \
        stmt;                             \
        (acc) += clock64() - _t0;         \
    } while (0)

struct PipeState {
    uint32_t idx = 0, phase = 0, n = kStages;
    __device__ __forceinline__ void advance() {
        if (++idx == n) {
            idx = 0;
            phase ^= 1;
        }
    }
};

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
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("nb200: cluster mbarrier watchdog block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}

template <int kPairs, typename hT, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_nvfp4_cluster_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_wgt,
                          const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_sfa,
                          const __grid_constant__ CUtensorMap tm_sfb, const __grid_constant__ CUtensorMap tm_lu, const ClusterParams p) {
    using Tr = HalfTraits<hT>;
    using S = SmemC;
    constexpr int kCtas = 2 * kPairs;
    extern __shared__ uint8_t smem_raw[];
    S &s = *reinterpret_cast<S *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

    const long long t_kernel0 = clock64();
    unsigned long long ns_kernel0 = 0;
    if (p.prof && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_kernel0));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pi = static_cast<int>(rank >> 1);      // pair inside the cluster
    const int q = static_cast<int>(rank & 1);        // rank inside the pair
    const bool leader = q == 0;
    const int cluster_id = blockIdx.x / kCtas;
    const int num_clusters = gridDim.x / kCtas;
    const int k64_total = p.K >> 6;
    const int num_kblocks = (k64_total + 3) >> 2;
    // NB200_GEMM_DEBUG ablation bits (results invalid): 4 no main-loop MMAs, 8 no epilogue math/stores, 16 no scale-factor
    // copies, 32 the MMA warp does not wait for operands, 64 the producer issues no main-loop loads (use with 32), 128 no low-rank,
    // 512 low-rank activations converted at the start of their own tile instead of one tile ahead (results stay valid).
    // Tried and dropped (r02, tools/gemm_ablate.py on one box): scale factors on their own mbarrier, loaded first and copied to TMEM one
    // stage ahead -- 12 % SLOWER on 4096x3072x3072, 26 % slower on K = 12288 (34.8 -> 39.0 us, 86 -> 108 us)
    const int lora_chunks = (p.has_lora && !(p.debug & 128)) ? p.Rp / kLoraChunk : 0;
    // one 32-rank chunk: the low-rank activations of tile it + 1 are converted while tile it is still in its main loop, into the other
    // lora_a buffer (with more chunks the single buffer 0 is recycled chunk by chunk inside a tile: lora_empty hand-shake)
    const bool lora_ahead = !(p.debug & 512) && lora_chunks <= 1;
    const uint32_t ring = (p.debug & 2048) ? 2 : ((p.debug & 1024) ? 3 : kStages);   // ablation: shallower TMA ring (results stay valid)
    constexpr uint16_t kMaskAll = (1u << kCtas) - 1;
    const uint16_t mask_pair = static_cast<uint16_t>(3u << (2 * pi));
    const uint16_t mask_a = static_cast<uint16_t>(kPairs == 2 ? ((1u << q) | (1u << (q + 2))) : (1u << q));

    if (warp == 0 && elect_one()) {
        prefetch_tensormap(&tm_act);
        prefetch_tensormap(&tm_wgt);
        prefetch_tensormap(&tm_sfa);
        prefetch_tensormap(&tm_sfb);
        if (p.has_lora) prefetch_tensormap(&tm_lu);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < kStages; i++) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], kPairs);
        }
        mbar_init(&s.tmem_full, 1);
        mbar_init(&s.tmem_empty, 2 * kEpiWarps);
        mbar_init(&s.lora_b_full, 1);
        mbar_init(&s.lora_a_full[0], 2 * kEpiWarps);
        mbar_init(&s.lora_a_full[1], 2 * kEpiWarps);
        mbar_init(&s.lora_empty, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_cg2<512>(&s.tmem_base);
    tc_fence_before_sync();
    cluster_sync();
    tc_fence_after_sync();
    const uint32_t tmem_base = s.tmem_base;
    const long long t_setup = clock64() - t_kernel0;
    griddep_launch_dependents();   // the next kernel in the stream may start its own setup on SMs this grid leaves idle
    griddep_wait();                // ... and this one touches global memory only after its predecessor has completed

    if (warp == 0) {
        // =================================== TMA producer (every CTA) ================================
        // the whole warp runs the loops and polls the barriers (uniform control flow, state in uniform registers); one elected lane issues
        {
            PipeState st;
            st.n = ring;
            uint32_t lora_phase = 0;
            long long t_empty = 0;
            for (int ct = cluster_id; ct < p.num_ct; ct += num_clusters) {
                const int mb2 = ct / p.nct_n, nb = (ct % p.nct_n) * kPairs + pi;
                const int m0 = mb2 * 2 * BM + q * BM;      // this CTA's A rows
                const int n0 = nb * BN;
                const int nh = n0 + q * BH;                // this CTA's B rows
                const int sfa_row = (m0 / 128) * k64_total, sfb_row = (n0 / 128 + q) * k64_total;
                for (int kb = 0; kb < num_kblocks; kb++) {
                    if (p.debug & 64) break;
                    if (p.prof) NB200_TIMED(t_empty, mbar_wait(&s.empty[st.idx], st.phase ^ 1));
                    else mbar_wait(&s.empty[st.idx], st.phase ^ 1);
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(&s.full[st.idx], 2 * kStageBytes);
                        if constexpr (kPairs == 2) {
                            // half of the A rows / activation scale blocks, multicast to the CTA with the same q in the other pair
                            tma_load_2d_cg2_mc(s.a[st.idx] + pi * (kABytes / 2), &tm_act, &s.full[st.idx], kb * 128, m0 + pi * (BM / 2), mask_a);
                            tma_load_2d_cg2_mc(s.sa[st.idx] + pi * (kSaBytes / 2), &tm_sfa, &s.full[st.idx], 0, sfa_row + 4 * kb + 2 * pi, mask_a);
                        } else {
                            tma_load_2d_cg2(s.a[st.idx], &tm_act, &s.full[st.idx], kb * 128, m0);
                            tma_load_2d_cg2(s.sa[st.idx], &tm_sfa, &s.full[st.idx], 0, sfa_row + 4 * kb);
                        }
                        tma_load_2d_cg2(s.b[st.idx], &tm_wgt, &s.full[st.idx], kb * 128, nh);
                        if constexpr (kPairs == 2) {
                            // weight scale factors of columns [n0 + 128 q, +128), multicast to both CTAs of the pair
                            tma_load_2d_cg2_mc(s.sb[st.idx] + q * (kSbBytes / 2), &tm_sfb, &s.full[st.idx], 0, sfb_row + 4 * kb, mask_pair);
                        } else {
                            // one pair per cluster: each CTA fetches the scale factors of all 256 columns itself -- measured ~4 % faster than
                            // splitting them and multicasting inside the pair (tools/gemm_ablate.py, r02)
                            tma_load_2d_cg2(s.sb[st.idx], &tm_sfb, &s.full[st.idx], 0, (n0 / 128) * k64_total + 4 * kb);
                            tma_load_2d_cg2(s.sb[st.idx] + kSbBytes / 2, &tm_sfb, &s.full[st.idx], 0, (n0 / 128 + 1) * k64_total + 4 * kb);
                        }
                    }
                    __syncwarp();
                    st.advance();
                }
                for (int c = 0; c < lora_chunks; c++) {
                    mbar_wait(&s.lora_empty, lora_phase ^ 1);
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(&s.lora_b_full, 2 * BH * kLoraChunk * 2);
                        // lora_up blocks [Rp/32][N/8][4][8][8] viewed as rows of 256 hT: row = chunk * N/8 + n/8
                        tma_load_2d_cg2(s.lora_b, &tm_lu, &s.lora_b_full, 0, c * (p.N >> 3) + (nh >> 3));
                    }
                    __syncwarp();
                    lora_phase ^= 1;
                }
            }
            if (p.prof && lane == 0) p.prof[blockIdx.x * 16 + 0] = t_empty;
        }
    } else if (warp == 1 && leader && (p.debug & ~12288) == 0 && (p.prof == nullptr || (p.debug & 4096))) {
        // ==================================== MMA issuer (pair leaders), hot path ======================
        // ONE thread feeds the tensor pipe: 12 tcgen05.cp + 4 tcgen05.mma + 1 commit per 512 clk of MMA time, so every cycle it spends
        // on anything else starves the pipe (timeline of tools/gemm_prof.py, r02a: 618 clk to issue a stage + 290 clk in an already
        // satisfied barrier wait = the 930 clk stage period).  Here the WHOLE WARP runs the loops and polls the barriers (uniform control
        // flow: loop state, descriptors and TMEM addresses live in uniform registers, no R2UR in front of every tcgen05 operand) and one
        // elected lane issues; descriptors are base + slot * stride; no clock reads, no debug branches.
        PipeState st;
        uint32_t lora_phase = 0, acc_phase = 0;
        constexpr uint32_t idesc_main = make_idesc_nvf4(2 * BM, BN);
        constexpr uint32_t idesc_lora = make_idesc_f16(Tr::kIsBf16, 2 * BM, BN);
        const uint64_t adesc0 = make_sw128_kmajor_desc(smem_u32(s.a[0])), bdesc0 = make_sw128_kmajor_desc(smem_u32(s.b[0]));
        const uint64_t sadesc0 = make_smem_desc(smem_u32(s.sa[0]), 0, 128, kLayoutNoSwizzle);
        const uint64_t sbdesc0 = make_smem_desc(smem_u32(s.sb[0]), 0, 128, kLayoutNoSwizzle);
        const uint64_t ladesc0 = make_smem_desc(smem_u32(s.lora_a[0]), 128, 512, kLayoutNoSwizzle);
        constexpr uint64_t kLoraABufDesc = (BM * kLoraChunk * 2) >> 4;
        uint32_t it = 0;
        const uint64_t lbdesc0 = make_smem_desc(smem_u32(s.lora_b), 128, 512, kLayoutNoSwizzle);
        const uint32_t tmem_d = tmem_base;
        // NB200_GEMM_DEBUG=4096 + a prof buffer: timeline of block 0's issuing lane (tools/gemm_prof.py --timeline), otherwise untouched
        long long *tl = ((p.debug & 4096) && p.prof != nullptr && blockIdx.x == 0) ? p.prof + 160 * 16 : nullptr;
        int tli = 0;
        auto mark = [&](int tag) {
            if (tl != nullptr && tli < 255) tl[tli++] = ((clock64() - t_kernel0) << 4) | tag;   // same origin as the epilogue timeline
        };
        const bool cluster_acq = (p.debug & 8192) != 0;
        for (int ct = cluster_id; ct < p.num_ct; ct += num_clusters) {
            if (lane == 0) mark(1);
            if (cluster_acq) mbar_wait_cluster(&s.tmem_empty, acc_phase ^ 1);
            else mbar_wait(&s.tmem_empty, acc_phase ^ 1);
            if (lane == 0) mark(2);
            tc_fence_after_sync();
            for (int kb = 0; kb < num_kblocks; kb++) {
                const int nj = min(4, k64_total - 4 * kb);
                const uint32_t sf_set = tmem_base + (kb & 1) * kSfSet;   // two scale-factor sets in TMEM (k-block parity)
                // CTA-scope acquire on every wait of this warp: what it orders is async-proxy work (TMA writes, tcgen05 reads) and the
                // peer's fence.proxy.async precedes its arrive, the same contract CUTLASS's ClusterBarrier::wait relies on; a satisfied
                // cluster-scope try_wait costs ~250 clk on the issuing thread every stage (tools/gemm_prof.py --timeline); bit 8192 = old way
                if (cluster_acq) mbar_wait_cluster(&s.full[st.idx], st.phase);
                else mbar_wait(&s.full[st.idx], st.phase);
                if (lane == 0) mark(3);
                tc_fence_after_sync();
                if (elect_one()) {
                    const uint64_t ad = adesc0 + static_cast<uint64_t>(st.idx) * (kABytes >> 4), bd = bdesc0 + static_cast<uint64_t>(st.idx) * (kBBytes >> 4);
                    const uint64_t sad = sadesc0 + static_cast<uint64_t>(st.idx) * (kSaBytes >> 4), sbd = sbdesc0 + static_cast<uint64_t>(st.idx) * (kSbBytes >> 4);
                    if (nj == 4) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            tc_cp_32x128b_warpx4_cg2(sf_set + kTmemSfa + 4 * j, sad + 32 * j);
                            tc_cp_32x128b_warpx4_cg2(sf_set + kTmemSfa + 16 + 8 * j, sbd + 32 * j);
                            tc_cp_32x128b_warpx4_cg2(sf_set + kTmemSfa + 16 + 8 * j + 4, sbd + 128 + 32 * j);
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            tc_mma_nvf4_cg2(tmem_d, ad + 2 * j, bd + 2 * j, idesc_main, sf_set + kTmemSfa + 4 * j, sf_set + kTmemSfa + 16 + 8 * j, (kb | j) != 0);
                    } else {   // K tail
                        for (int j = 0; j < nj; j++) {
                            tc_cp_32x128b_warpx4_cg2(sf_set + kTmemSfa + 4 * j, sad + 32 * j);
                            tc_cp_32x128b_warpx4_cg2(sf_set + kTmemSfa + 16 + 8 * j, sbd + 32 * j);
                            tc_cp_32x128b_warpx4_cg2(sf_set + kTmemSfa + 16 + 8 * j + 4, sbd + 128 + 32 * j);
                        }
                        for (int j = 0; j < nj; j++)
                            tc_mma_nvf4_cg2(tmem_d, ad + 2 * j, bd + 2 * j, idesc_main, sf_set + kTmemSfa + 4 * j, sf_set + kTmemSfa + 16 + 8 * j, (kb | j) != 0);
                    }
                    tc_commit_cg2(&s.empty[st.idx], kMaskAll);
                }
                __syncwarp();
                if (lane == 0) mark(4);
                st.advance();
            }
            for (int c = 0; c < lora_chunks; c++) {
                mbar_wait(&s.lora_b_full, lora_phase);
                const uint32_t lb = lora_ahead ? (it & 1) : 0, lph = lora_ahead ? ((it >> 1) & 1) : lora_phase;
                if (cluster_acq) mbar_wait_cluster(&s.lora_a_full[lb], lph);
                else mbar_wait(&s.lora_a_full[lb], lph);
                tc_fence_after_sync();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < kLoraChunk / 16; j++) tc_mma_f16_cg2(tmem_d, ladesc0 + lb * kLoraABufDesc + 16 * j, lbdesc0 + 16 * j, idesc_lora, 1);
                    tc_commit_cg2(&s.lora_empty, mask_pair);
                }
                __syncwarp();
                lora_phase ^= 1;
            }
            if (elect_one()) tc_commit_cg2(&s.tmem_full, mask_pair);
            __syncwarp();
            acc_phase ^= 1;
            it++;
        }
    } else if (warp == 1) {
        // ============ MMA issuer, instrumented path (profiling counters / ablation bits): one thread, clock reads around the waits =======
        if (leader && elect_one()) {
            PipeState st;
            st.n = ring;
            uint32_t lora_phase = 0, acc_phase = 0;
            long long t_tmem_empty = 0, t_full = 0, t_lora = 0, t_first = 0;
            const long long t_mma0 = clock64();
            constexpr uint32_t idesc_main = make_idesc_nvf4(2 * BM, BN);
            constexpr uint32_t idesc_lora = make_idesc_f16(Tr::kIsBf16, 2 * BM, BN);
            constexpr bool lean = false;
            // timeline of block 0 (tools/gemm_prof.py --timeline): clock after each wait, 256 slots behind the per-CTA counters
            long long *tl = (p.prof && blockIdx.x == 0) ? p.prof + 160 * 16 : nullptr;
            int tli = 0;
            auto mark = [&](int tag) {
                if (tl && tli < 255) tl[tli++] = ((clock64() - t_mma0) << 4) | tag;
            };
            uint32_t it = 0;
            for (int ct = cluster_id; ct < p.num_ct; ct += num_clusters, it++) {
                mark(1);
                if (lean) mbar_wait_cluster(&s.tmem_empty, acc_phase ^ 1);
                else NB200_TIMED(t_tmem_empty, mbar_wait_cluster(&s.tmem_empty, acc_phase ^ 1));
                mark(2);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base;
                for (int kb = 0; kb < num_kblocks; kb++) {
                    const int nj = min(4, k64_total - 4 * kb);
                    // two scale-factor sets in TMEM (k-block parity): the copies of stage s+1 do not wait for the MMAs of stage s
                    const uint32_t sf_set = tmem_base + (kb & 1) * kSfSet;
                    if (!(p.debug & 32)) NB200_TIMED(t_full, mbar_wait_cluster(&s.full[st.idx], st.phase));
                    mark(3);
                    if (t_first == 0) t_first = clock64() - t_mma0;
                    tc_fence_after_sync();
                    for (int j = 0; j < ((p.debug & 16) ? 0 : nj); j++) {
                        tc_cp_32x128b_warpx4_cg2(sf_set + kTmemSfa + 4 * j, make_smem_desc(smem_u32(s.sa[st.idx] + j * 512), 0, 128, kLayoutNoSwizzle));
#pragma unroll
                        for (int h = 0; h < 2; h++)
                            tc_cp_32x128b_warpx4_cg2(sf_set + kTmemSfa + 16 + 8 * j + 4 * h,
                                                     make_smem_desc(smem_u32(s.sb[st.idx] + h * 2048 + j * 512), 0, 128, kLayoutNoSwizzle));
                    }
                    const uint32_t a_addr = smem_u32(s.a[st.idx]), b_addr = smem_u32(s.b[st.idx]);
                    for (int j = 0; j < ((p.debug & 4) ? 0 : nj); j++)
                        tc_mma_nvf4_cg2(tmem_d, make_sw128_kmajor_desc(a_addr + j * 32), make_sw128_kmajor_desc(b_addr + j * 32), idesc_main,
                                        sf_set + kTmemSfa + 4 * j, sf_set + kTmemSfa + 16 + 8 * j, (kb | j) != 0);
                    tc_commit_cg2(&s.empty[st.idx], kMaskAll);
                    mark(4);
                    st.advance();
                }
                for (int c = 0; c < lora_chunks; c++) {
                    const uint32_t lb = lora_ahead ? (it & 1) : 0, lph = lora_ahead ? ((it >> 1) & 1) : lora_phase;
                    if (lean) {
                        mbar_wait_cluster(&s.lora_b_full, lora_phase);
                        mbar_wait_cluster(&s.lora_a_full[lb], lph);
                    } else {
                        NB200_TIMED(t_lora, mbar_wait_cluster(&s.lora_b_full, lora_phase));
                        NB200_TIMED(t_lora, mbar_wait_cluster(&s.lora_a_full[lb], lph));
                    }
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(s.lora_a[lb]), b_addr = smem_u32(s.lora_b);
#pragma unroll
                    for (int j = 0; j < kLoraChunk / 16; j++)
                        tc_mma_f16_cg2(tmem_d, make_smem_desc(a_addr + j * 256, 128, 512, kLayoutNoSwizzle),
                                       make_smem_desc(b_addr + j * 256, 128, 512, kLayoutNoSwizzle), idesc_lora, 1);
                    tc_commit_cg2(&s.lora_empty, mask_pair);
                    lora_phase ^= 1;
                }
                tc_commit_cg2(&s.tmem_full, mask_pair);
                acc_phase ^= 1;
            }
            if (p.prof) {
                p.prof[blockIdx.x * 16 + 1] = t_tmem_empty;
                p.prof[blockIdx.x * 16 + 2] = t_full;
                p.prof[blockIdx.x * 16 + 3] = t_lora;
                p.prof[blockIdx.x * 16 + 10] = clock64() - t_mma0;
                p.prof[blockIdx.x * 16 + 13] = t_first;
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================================== epilogue (every CTA) ==================================
        constexpr int H = kEpiGroups;
        const int h = (warp - kEpiWarp0) >> 2;        // column half: [128 h, +128) of the tile
        const int qd = warp & 3;                      // TMEM lane quadrant
        const int et = threadIdx.x - (kEpiWarp0 + 4 * h) * 32;
        const int eta = threadIdx.x - kEpiWarp0 * 32;
        const int row = qd * 32 + lane;
        uint32_t lora_phase = 0, acc_phase = 0;
        long long t_tmem_full = 0, t_pre = 0;
        const long long t_epi0 = clock64();
        const uint32_t leader_rank = rank & ~1u;
        const uint32_t lora_a_full_leader[2] = {mapa(smem_u32(&s.lora_a_full[0]), leader_rank), mapa(smem_u32(&s.lora_a_full[1]), leader_rank)};
        const uint32_t tmem_empty_leader = mapa(smem_u32(&s.tmem_empty), leader_rank);
        if constexpr (EPI == EPI_ROPE) {
            for (int i = eta; i < 256; i += kEpiThreads)
                s.normw[i] = Tr::to_float(reinterpret_cast<const hT *>(i < 128 ? p.norm_q : p.norm_k)[i & 127]);
        }
        [[maybe_unused]] const bool pack_qkv = EPI == EPI_ROPE && p.out_qkv[0] != nullptr;
        // low-rank activations of one tile: fp32 -> * lora_scale -> hT (lora.cuh:145-151) into the UMMA-layout A operand; one row per
        // thread, group h converts 16 of the 32 ranks of a chunk.  Called ONE TILE AHEAD (right after this tile's accumulator has
        // been pulled into registers): the MMA warp reaches the low-rank MMAs of tile i+1 while these warps are still busy with the
        // math and stores of tile i, and must not wait for them.
        // NB200_GEMM_DEBUG=4096 + a prof buffer of >= 192 rows: timeline of block 0's first epilogue thread (tools/gemm_prof.py --timeline)
        long long *etl = ((p.debug & 4096) && p.prof != nullptr && blockIdx.x == 0 && eta == 0) ? p.prof + 176 * 16 : nullptr;
        int etli = 0;
        auto emark = [&](int tag) {
            if (etl != nullptr && etli < 255) etl[etli++] = ((clock64() - t_kernel0) << 4) | tag;
        };
        auto convert_lora = [&](const int m0_, const uint32_t buf) {
            for (int c = 0; c < lora_chunks; c++) {
                if (!lora_ahead) mbar_wait(&s.lora_empty, lora_phase ^ 1);   // (ahead: the other buffer's last reader finished a tile ago)
                const float *src = p.lora_act + static_cast<size_t>(m0_ + row) * p.R + c * kLoraChunk;
                uint8_t *dst = s.lora_a[buf] + (row >> 3) * 512 + (row & 7) * 16;
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
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(lora_a_full_leader[buf]);
                lora_phase ^= 1;
            }
        };
        // bias / per-channel scale of one tile -> shared memory (thread eta owns column eta: a group only reads what its own threads wrote)
        auto fetch_channel_vectors = [&](const int n0_, const uint32_t buf) {
            for (int i = eta; i < BN; i += kEpiThreads) {
                s.bias[buf][i] = p.bias != nullptr ? p.bias[n0_ + i] : 0.f;
                s.cscale[buf][i] = p.cscale != nullptr ? p.cscale[n0_ + i] : 1.f;
            }
        };
        auto tile_m0 = [&](const int ct_) { return (ct_ / p.nct_n) * 2 * BM + q * BM; };
        auto tile_n0 = [&](const int ct_) { return ((ct_ % p.nct_n) * kPairs + pi) * BN; };
        if (cluster_id < p.num_ct) {
            fetch_channel_vectors(tile_n0(cluster_id), 0);
            if (lora_ahead) convert_lora(tile_m0(cluster_id), 0);
        }
        uint32_t it = 0;
        for (int ct = cluster_id; ct < p.num_ct; ct += num_clusters, it++) {
            const int m0 = tile_m0(ct), n0 = tile_n0(ct);
            const long long t_tile0 = clock64();
            const float *bias_s = s.bias[it & 1], *cscale_s = s.cscale[it & 1];
            // While this tile's MMAs are still running: everything the NEXT tile's epilogue (and its low-rank MMAs) will need.  The
            // barrier orders the previous iteration's writes of this tile's vectors before their use below, and the previous tile's
            // reads of the other buffer before its refill here.
            named_bar_sync(2 + 2 * h, 128);
            if (ct + num_clusters < p.num_ct) {
                fetch_channel_vectors(tile_n0(ct + num_clusters), (it + 1) & 1);
                if (lora_ahead) convert_lora(tile_m0(ct + num_clusters), (it + 1) & 1);
            }
            if (!lora_ahead) convert_lora(m0, 0);

            t_pre += clock64() - t_tile0;
            emark(1);
            NB200_TIMED(t_tmem_full, mbar_wait(&s.tmem_full, acc_phase));
            emark(2);
            acc_phase ^= 1;
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + h * 128;
            // this group's 128 columns -> registers, then hand the accumulator back before any math
            uint32_t va[32], vb[32], vc[32], vd[32];
            tmem_ld_32x32b_x32(taddr, va);
            tmem_ld_32x32b_x32(taddr + 32, vb);
            tmem_ld_32x32b_x32(taddr + 64, vc);
            tmem_ld_32x32b_x32(taddr + 96, vd);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tmem_empty_leader);
            emark(3);
            if (p.debug & 8) continue;

            // ---- EPI_ROPE: one 128-wide head per group
            [[maybe_unused]] bool do_rope = false;
            [[maybe_unused]] int qkv_part = 0, qkv_head = 0;
            [[maybe_unused]] float rope_coef = 1.f;
            [[maybe_unused]] const float *normw = s.normw;
            [[maybe_unused]] const float *rot_row = nullptr;
            if constexpr (EPI == EPI_ROPE) {
                const int third = p.N / 3;
                const int col0 = n0 + h * 128;
                qkv_part = col0 / third;       // 0 = Q heads, 1 = K heads, 2 = V (untouched)
                qkv_head = (col0 % third) >> 7;
                do_rope = qkv_part < 2;
                if (do_rope) {
                    // sum of squares of the head's 128 columns: 8 partial sums of 16 columns (sequential FMA chains) combined by the fixed tree
                    // ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)) -- the order csrc/rope.cu's 8-lanes-per-head reduction produces (bit-identical routes)
                    auto sq32 = [&](const uint32_t(&v)[32], const int cbase) {   // -> p_even + p_odd of 32 columns
                        float pa = 0.f, pb = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 cs = *reinterpret_cast<const float4 *>(cscale_s + cbase + i), bs = *reinterpret_cast<const float4 *>(bias_s + cbase + i);
                            const float y0 = fmaf(__uint_as_float(v[i]), cs.x, bs.x), y1 = fmaf(__uint_as_float(v[i + 1]), cs.y, bs.y);
                            const float y2 = fmaf(__uint_as_float(v[i + 2]), cs.z, bs.z), y3 = fmaf(__uint_as_float(v[i + 3]), cs.w, bs.w);
                            const float2 r = Tr::to_float2(Tr::from_float2(make_float2(y0, y1)));   // fpsum is hT (epilogues.cuh:327-341)
                            const float2 r2 = Tr::to_float2(Tr::from_float2(make_float2(y2, y3)));
                            if (i < 16) {
                                pa = fmaf(r.x, r.x, pa);
                                pa = fmaf(r.y, r.y, pa);
                                pa = fmaf(r2.x, r2.x, pa);
                                pa = fmaf(r2.y, r2.y, pa);
                            } else {
                                pb = fmaf(r.x, r.x, pb);
                                pb = fmaf(r.y, r.y, pb);
                                pb = fmaf(r2.x, r2.x, pb);
                                pb = fmaf(r2.y, r2.y, pb);
                            }
                        }
                        return pa + pb;
                    };
                    const float q0 = sq32(va, h * 128) + sq32(vb, h * 128 + 32);
                    const float q1 = sq32(vc, h * 128 + 64) + sq32(vd, h * 128 + 96);
                    const float sumsq = q0 + q1;
                    rope_coef = rsqrt_approx_ftz(sumsq / 128.f + 1e-6f);
                    normw = s.normw + qkv_part * 128;
                    // reference pack_rotemb order (transformer_flux.py:60-92): float index of (row m, pair pr, sin|cos)
                    //   ((((m/16*16 + pr/4)*8 + m%8)*4 + pr%4)*2 + (m%16)/8)*2 + {0,1}
                    const int m = m0 + row;
                    rot_row = p.rotary + (static_cast<size_t>(m >> 4) * 16 * 8 + (m & 7)) * 16 + ((m >> 3) & 1) * 2;
                }
            }

            // one 64-column chunk (cc = 0, 1 inside this group's half): scale / bias / activation / RoPE, hT pack, store
            auto do_chunk = [&](const int cc, const uint32_t(&v0)[32], const uint32_t(&v1)[32]) {
                const int ch = h * 2 + cc;   // chunk inside the 256-wide tile
                // Output tile: each warp transposes its own 32 rows x 64 columns through its 4 KB of staging (row per lane in, 4 rows x 128
                // contiguous bytes per store instruction out).  No TMA store: it queues behind the operand loads in the SM's one TMA unit
                // (r02 epilogue timeline: 1.4k clk until a 16 KB store had left shared memory) and needs group-wide barriers on both sides.
                __syncwarp();   // this warp's read-out of the previous chunk is done
                emark(5);
                uint8_t *srow = s.out_stage[h] + row * 128;
                [[maybe_unused]] __half *qkv_row = nullptr;
                [[maybe_unused]] bool qkv_masked = false;
                if constexpr (EPI == EPI_ROPE) {
                    if (pack_qkv) {
                        qkv_row = p.out_qkv[qkv_part] + qkv_head * p.stride_head[qkv_part] + static_cast<size_t>(m0 + row) * 128 + cc * 64;
                        qkv_masked = m0 + row >= p.attn_tokens;
                    }
                }
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) {
                    uint32_t w[4];
                    const float4 csA = *reinterpret_cast<const float4 *>(cscale_s + ch * 64 + c8 * 8), csB = *reinterpret_cast<const float4 *>(cscale_s + ch * 64 + c8 * 8 + 4);
                    const float4 bsA = *reinterpret_cast<const float4 *>(bias_s + ch * 64 + c8 * 8), bsB = *reinterpret_cast<const float4 *>(bias_s + ch * 64 + c8 * 8 + 4);
                    const float cs8[8] = {csA.x, csA.y, csA.z, csA.w, csB.x, csB.y, csB.z, csB.w};
                    const float bs8[8] = {bsA.x, bsA.y, bsA.z, bsA.w, bsB.x, bsB.y, bsB.z, bsB.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int col = c8 * 8 + 2 * i;
                        float y0 = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]);
                        float y1 = __uint_as_float(col + 1 < 32 ? v0[col + 1] : v1[col + 1 - 32]);
                        y0 = fmaf(y0, cs8[2 * i], bs8[2 * i]);
                        y1 = fmaf(y1, cs8[2 * i + 1], bs8[2 * i + 1]);
                        if constexpr (EPI == EPI_ROPE) {
                            if (do_rope) {
                                const float2 r = Tr::to_float2(Tr::from_float2(make_float2(y0, y1)));
                                const int hc = cc * 64 + col;   // column inside the head
                                const float x0 = r.x * (rope_coef * normw[hc]);
                                const float x1 = r.y * (rope_coef * normw[hc + 1]);
                                const int pr = hc >> 1;
                                const float2 sc = *reinterpret_cast<const float2 *>(rot_row + (pr >> 2) * 128 + (pr & 3) * 4);
                                y0 = x0 * sc.y - x1 * sc.x;   // (sin, cos) = (sc.x, sc.y)  epilogues.cuh:362-367
                                y1 = x0 * sc.x + x1 * sc.y;
                            }
                        } else {
                            if (p.mid_act != NB200_ACT_NONE) {   // the reference applies the activation to the hT-rounded value
                                const float2 r = Tr::to_float2(Tr::from_float2(make_float2(y0, y1)));
                                if (p.mid_act == NB200_ACT_GELU) {
                                    y0 = gelu_f32(r.x);
                                    y1 = gelu_f32(r.y);
                                } else {
                                    y0 = silu_f32(r.x);
                                    y1 = silu_f32(r.y);
                                }
                            }
                        }
                        if constexpr (!Tr::kIsBf16) {   // fp16 stores clamp (gemm_base.cuh:688-696)
                            y0 = fminf(fmaxf(y0, -65504.f), 65504.f);
                            y1 = fminf(fmaxf(y1, -65504.f), 65504.f);
                        }
                        typename Tr::T2 hv = Tr::from_float2(make_float2(y0, y1));
                        w[i] = *reinterpret_cast<uint32_t *>(&hv);
                        if constexpr (EPI == EPI_ROPE) {
                            if (pack_qkv) {   // hT -> fp16 through fp32 (epilogues.cuh:446-453); pad rows masked 0 / NaN / 0
                                const __half2 hh = __float22half2_rn(Tr::to_float2(hv));
                                w[i] = qkv_masked ? (qkv_part == 1 ? 0x7FFF7FFFu : 0u) : *reinterpret_cast<const uint32_t *>(&hh);
                            }
                        }
                    }
                    if constexpr (EPI == EPI_ROPE) {
                        if (pack_qkv) {
                            *reinterpret_cast<uint4 *>(qkv_row + c8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
                            continue;
                        }
                    }
                    *reinterpret_cast<uint4 *>(srow + ((c8 ^ (row & 7)) * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                if (pack_qkv) return;
                emark(6);
                __syncwarp();
                {
                    const int c = lane & 7;
                    const int col = n0 + ch * 64 + c * 8;
                    hT *gout = static_cast<hT *>(p.out);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int r = qd * 32 + j * 4 + (lane >> 3);
                        const uint4 v = *reinterpret_cast<const uint4 *>(s.out_stage[h] + r * 128 + ((c ^ (r & 7)) * 16));
                        if (m0 + r < p.M_out && col < p.N_out) *reinterpret_cast<uint4 *>(gout + static_cast<size_t>(m0 + r) * p.N_out + col) = v;
                    }
                }
                emark(7);
            };
            do_chunk(0, va, vb);
            do_chunk(1, vc, vd);
        }
        if (p.prof && eta == 0) {
            p.prof[blockIdx.x * 16 + 4] = t_tmem_full;
            p.prof[blockIdx.x * 16 + 5] = clock64() - t_epi0;
            p.prof[blockIdx.x * 16 + 11] = t_pre;
        }
    }

    // every CTA must stay alive until the cluster's MMAs, multicasts and remote arrivals are done
    tc_fence_before_sync();
    cluster_sync();
    if (p.prof && threadIdx.x == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        p.prof[blockIdx.x * 16 + 9] = clock64() - t_kernel0;
        p.prof[blockIdx.x * 16 + 12] = t_setup;
        p.prof[blockIdx.x * 16 + 14] = static_cast<long long>(ns1 - ns_kernel0);   // wall time of this CTA: effective SM clock = slot 9 / slot 14
    }
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc_cg2<512>(tmem_base);
    }
}

constexpr int kMaxDevices = 64;   // per-device launch state (one process may drive several GPUs)

template <int kPairs, typename hT, int EPI>
int launch_cluster(const nb200_gemm_args &a, cudaStream_t stream) {
    constexpr int kCtas = 2 * kPairs;
    CUtensorMap tm_act, tm_wgt, tm_out, tm_sfa, tm_sfb, tm_lu;
    int rc = make_map_2d(&tm_act, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.act, a.K / 2, a.Mp, a.K / 2, 128, kPairs == 2 ? BM / 2 : BM,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_2d(&tm_wgt, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.wgt, a.K / 2, a.N, a.K / 2, 128, BH, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const CUtensorMapDataType odt = HalfTraits<hT>::kIsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (a.out != nullptr) {
        rc = make_map_2d(&tm_out, odt, a.out, a.N_out, a.M_out, static_cast<uint64_t>(a.N_out) * 2, 64, BM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    } else {
        tm_out = tm_act;   // PackQKV: never dereferenced
    }
    const uint64_t k64 = a.K / 64;
    // scale-factor tiles as rows of 128 x u32 (512 B): row = (rows / 128) * K/64 + k64 block
    rc = make_map_2d(&tm_sfa, CU_TENSOR_MAP_DATA_TYPE_UINT32, a.ascales, 128, (a.Mp / 128) * k64, 512, 128, kPairs == 2 ? 2 : 4,
                     CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_map_2d(&tm_sfb, CU_TENSOR_MAP_DATA_TYPE_UINT32, a.wscales, 128, (a.N / 128) * k64, 512, 128, 4, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    tm_lu = tm_act;
    const int Rp = (a.R_up + 31) / 32 * 32;
    if (a.R_up > 0) {
        rc = make_map_2d(&tm_lu, odt, a.lora_up, 256, static_cast<uint64_t>(Rp / 32) * (a.N / 8), 512, 256, BH / 8, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    ClusterParams p;
    p.bias = a.bias;
    p.cscale = a.cscale;
    p.lora_act = a.lora_act_in;
    p.has_lora = a.R_up > 0;
    p.Mp = a.Mp;
    p.N = a.N;
    p.K = a.K;
    p.R = a.R_up;
    p.Rp = Rp;
    p.nct_n = a.N / (BN * kPairs);
    p.num_ct = (a.Mp / (2 * BM)) * p.nct_n;
    p.mid_act = a.mid_act;
    p.norm_q = a.norm_q;
    p.norm_k = a.norm_k;
    p.rotary = a.rotary_emb;
    p.out_qkv[0] = static_cast<__half *>(a.out_q);
    p.out_qkv[1] = static_cast<__half *>(a.out_k);
    p.out_qkv[2] = static_cast<__half *>(a.out_v);
    p.stride_head[0] = a.stride_head_q;
    p.stride_head[1] = a.stride_head_k;
    p.stride_head[2] = a.stride_head_v;
    p.attn_tokens = a.attn_tokens;
    p.out = a.out;
    p.M_out = a.M_out;
    p.N_out = a.N_out;
    p.prof = static_cast<long long *>(a.prof);
    const char *dbg_env = getenv("NB200_GEMM_DEBUG");   // read per launch: tools/gemm_ablate.py sweeps it inside one process
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    p.debug = dbg;
    for (int i = 0; i < NB200_MAX_LORA_SCALES; i++) p.lora_scales[i] = a.lora_scales[i];

    auto kern = gemm_nvfp4_cluster_kernel<kPairs, hT, EPI>;
    const size_t smem_bytes = sizeof(SmemC) + 1024;
    if (int rc2 = set_max_smem_once(reinterpret_cast<const void *>(kern), smem_bytes)) return rc2;
    // clusters that can be co-resident (a cluster lives inside one GPC: 148 SMs do not always hold 37 clusters of 4): per device
    static std::mutex mu;
    static int max_clusters_dev[kMaxDevices] = {};
    int dev = 0;
    NB200_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(NB200_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    int max_clusters;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (max_clusters_dev[dev] == 0) {
            int num_sms = 0;
            if (int rc2 = current_device_sms(&num_sms)) return rc2;
            LaunchCfg probe(dim3(num_sms / kCtas * kCtas), dim3(kThreads), smem_bytes, stream, kCtas);
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &probe.cfg) != cudaSuccess || n <= 0) {
                (void)cudaGetLastError();
                n = num_sms / kCtas;
            }
            max_clusters_dev[dev] = n;
        }
        max_clusters = max_clusters_dev[dev];
    }
    if (a.num_sms > 0) max_clusters = std::max(1, std::min(max_clusters, a.num_sms / kCtas));
    const int clusters = p.num_ct < max_clusters ? p.num_ct : max_clusters;
    LaunchCfg lc(dim3(clusters * kCtas), dim3(kThreads), smem_bytes, stream, kCtas);
    NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, tm_act, tm_wgt, tm_out, tm_sfa, tm_sfb, tm_lu, p));
    count_launch();
    return NB200_OK;
}

template <typename hT>
int dispatch_t(const nb200_gemm_args &a, cudaStream_t stream, int pairs) {
    bool rope = a.rotary_emb != nullptr;
    // QKV with a plain [M, N] output: the GEMM with its default epilogue followed by the in-place RMSNorm + RoPE kernel (rope.cu) beats
    // the fused epilogue by a wide margin on these tile shapes (115.6 -> ~80 us at 4352 x 3072 -> 9216: the rotary table is 128 KB of
    // extra ingest per tile when every epilogue thread fetches its own row).  Bit-identical results; NB200_ROPE_SPLIT=0 keeps it fused.
    // PackQKV (three fp16 outputs) stays fused: it has no [M, N] buffer to work in.
    const char *split_env = getenv("NB200_ROPE_SPLIT");   // read per launch (tests and tools flip it inside one process)
    const bool split_enabled = !(split_env && atoi(split_env) == 0);
    if (rope && a.out != nullptr && a.out_q == nullptr && a.N_out == a.N && split_enabled) {
        nb200_gemm_args plain = a;
        plain.rotary_emb = nullptr;
        plain.norm_q = plain.norm_k = nullptr;
        const int rc = pairs == 2 ? launch_cluster<2, hT, EPI_DEFAULT>(plain, stream) : launch_cluster<1, hT, EPI_DEFAULT>(plain, stream);
        if (rc != NB200_OK) return rc;
        return rope_inplace_dispatch(a.dtype, a.out, a.M_out, a.N, a.norm_q, a.norm_k, a.rotary_emb, stream);
    }
    // PackQKV with a caller-provided [Mp, N] scratch in `out`: the same split -- plain GEMM into the scratch, then the RMSNorm + RoPE + pack kernel
    // writes the attention operands (bit-identical to the fused epilogue, tests/test_gpu_fused.py)
    if (rope && a.out != nullptr && a.out_q != nullptr && split_enabled) {
        nb200_gemm_args plain = a;
        plain.rotary_emb = nullptr;
        plain.norm_q = plain.norm_k = nullptr;
        plain.out_q = plain.out_k = plain.out_v = nullptr;
        const int rc = pairs == 2 ? launch_cluster<2, hT, EPI_DEFAULT>(plain, stream) : launch_cluster<1, hT, EPI_DEFAULT>(plain, stream);
        if (rc != NB200_OK) return rc;
        return rope_pack_dispatch(a.dtype, a.out, a.Mp, a.N, a.norm_q, a.norm_k, a.rotary_emb, a.out_q, a.out_k, a.out_v, a.stride_head_q, a.stride_head_k,
                                  a.stride_head_v, a.attn_tokens, stream);
    }
    if (rope && a.out != nullptr && a.out_q != nullptr) {   // split switched off: the scratch is not needed
        nb200_gemm_args fused = a;
        fused.out = nullptr;
        return pairs == 2 ? launch_cluster<2, hT, EPI_ROPE>(fused, stream) : launch_cluster<1, hT, EPI_ROPE>(fused, stream);
    }
    if (pairs == 2) return rope ? launch_cluster<2, hT, EPI_ROPE>(a, stream) : launch_cluster<2, hT, EPI_DEFAULT>(a, stream);
    return rope ? launch_cluster<1, hT, EPI_ROPE>(a, stream) : launch_cluster<1, hT, EPI_DEFAULT>(a, stream);
}

}  // namespace

// NVFP4, default / RMSNorm+RoPE epilogues, N % 256 == 0.  pairs = 0: two pairs per cluster (A multicast) when N % 512 == 0.
int gemm_nvfp4_cluster_dispatch(const nb200_gemm_args &a, cudaStream_t stream, int pairs) {
    if (!a.fp4 || a.N % BN != 0 || a.qout != nullptr) return fail(NB200_ERR_INVALID_ARGUMENT, "cluster kernel: NVFP4, N % 256 == 0, no fused quantise epilogue");
    if (pairs == 0) pairs = a.N % (2 * BN) == 0 ? 2 : 1;
    if (pairs == 2 && a.N % (2 * BN) != 0) return fail(NB200_ERR_INVALID_ARGUMENT, "two pairs per cluster need N % 512 == 0");
    return a.dtype == NB200_BF16 ? dispatch_t<__nv_bfloat16>(a, stream, pairs) : dispatch_t<__half>(a, stream, pairs);
}

}  // namespace nb200
