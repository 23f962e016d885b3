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

int gemm_w4a4_2cta_dispatch(const nb200_gemm_args &a, cudaStream_t stream);
int gemm_nvfp4_cluster_dispatch(const nb200_gemm_args &a, cudaStream_t stream, int pairs);

namespace {

using namespace ptx;

constexpr int BM = 128;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiThreads = 128;
constexpr int kConvWarp0 = 8;
// INT4: 24 converter warps -> one 32-element unit per thread per k block (128x64 A + 256x64 B = 768 units).
// Measured with 8 warps (3 units per thread): the converters were busy 75 % of the kernel and the MMA warp
// starved 85 % of it; the work is latency bound (LDS -> LOP3 -> HSUB2 -> HMUL2 -> STS chains), so it is
// spread over more warps and the register file is re-balanced with setmaxnreg (1024 threads x 64 regs).
constexpr int kConvWarps = 24;
constexpr int kNumConvThreads = kConvWarps * 32;
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
    // EPI_QUANT: fused GELU -> low-rank down projection (next layer) -> 4-bit quantise (next layer)
    void *out;               // non-null when the hT tile is stored too
    uint8_t *qout;           // [Mp][N/2]
    void *oscales_out;       // INT4 hT [N/64][Mp] | FP4 scale tiles
    const void *smooth_next; // hT [N]
    float *lora_act_out;     // [Mp][R_down], pre-zeroed
    int R_down, Rdp;
    int tile_contig;         // 1: each CTA owns a contiguous range of tiles (same m-block runs), see EPI_QUANT
    float *ws_partial;       // EPI_QUANT, optional: [m-block + CTA][128 rows][Rdp] partial down projections (deterministic reduction) or null
    // EPI_ROPE
    const void *norm_q, *norm_k;  // hT [128]
    const float *rotary;          // reference pack_rotemb layout [Mp][128]
    __half *out_qkv[3];           // EpiloguePackQKV: fp16 [heads][rows][128] each (null: store `out`)
    long long stride_head[3];
    int attn_tokens;
    // EPI_LITELA (SANA linear attention, epilogues.cuh:552-691): `out` is relu(Q) [Mp, N/3]; the K | V tiles are reduced over tokens into out_vk
    float *out_vk;                // f32 [batch][heads][33][32], zero before the launch, accumulated with fp32 atomics (as the reference's reduce_add)
    int vk_tokens, vk_heads;      // tokens per batch (multiple of 128: a tile never straddles two images), heads = N / 96
    long long *prof;              // optional [grid][16] cycle counters (tools/gemm_prof.py)
    int debug;                    // NB200_GEMM_DEBUG experiment bits (results invalid when non-zero): 1 = converters
                                  // skip their smem stores, 2 = converters skip all work, 4 = no main-loop MMAs, 8 = epilogue drains TMEM but skips math/stores,
                                  // 16 = EPI_QUANT: skip the lora_act_out atomics, 32 = EPI_QUANT: skip the next-layer quantise math
    float lora_scales[NB200_MAX_LORA_SCALES];
};

enum { EPI_DEFAULT = 0, EPI_QUANT = 1, EPI_ROPE = 2, EPI_LITELA = 3 };
constexpr int kMaxRdp = 128;  // largest fused next-layer rank (smem / TMEM budget)
constexpr int kWsMaxCtas = 256;  // runs the deterministic-reduction workspace is sized for (one CTA per SM)

template <bool FP4, int BN, int EPI = EPI_DEFAULT>
struct Cfg {
    // NVFP4 k-stage = kK64 blocks of 64 elements.  kK64 = 2 (64-byte rows, SWIZZLE_64B, a ring twice as deep) is implemented
    // and was measured on the plain 256-wide tile: SLOWER (M=4352 K=3072 N=12288: 120.7 vs 100.4 us; load pipeline alone 91 vs
    // 71 us) -- TMA throughput drops with the row length (32-byte rows ~16, 64-byte ~38, 128-byte ~49 B/clk/SM), so the
    // full 128-byte rows stay.
    static constexpr int kK64 = FP4 ? 4 : 1;
    static constexpr int kBK = 64 * kK64;                       // k elements per pipeline stage
    static constexpr int kStages = FP4 ? (kK64 == 2 ? 6 : ((BN == 256 || EPI == EPI_QUANT) ? 3 : 4)) : 4;  // TMA ring depth
    static constexpr int kConvStages = 2;                      // INT4: converted-tile ring depth
    // INT4 fused-quantise tiles of 256 columns keep ONE accumulator: 256 + the D2 columns must fit TMEM, and an INT4
    // main loop (~48 k-blocks x ~1000 clk) dwarfs the exposed epilogue
    static constexpr int kNumAcc = FP4 ? (BN <= 128 ? 2 : 1) : ((EPI == EPI_QUANT && BN == 256) ? 1 : 2);
    static constexpr int kMaxRdp = BN == 256 ? 32 : nb200::kMaxRdp;   // fused next-layer rank supported by this tile shape
    static constexpr int kABytes = BM * 32 * kK64;              // packed A tile per stage (32 bytes per row and K64 block)
    static constexpr int kBBytes = BN * 32 * kK64;
    static constexpr int kSfaCols = 16;                         // 4 K64 blocks x 4 columns
    static constexpr int kSfbCols = BN / 8;                     // 4 K64 blocks x BN/32 columns
    static constexpr int kTmemSfa = kNumAcc * BN;
    static constexpr int kTmemSfb = kTmemSfa + kSfaCols;
    // scale factors are DOUBLE BUFFERED in TMEM (set = k-stage parity): with one set the tcgen05.cp of stage s+1 has a
    // write-after-read hazard on the columns the MMAs of stage s are reading, so copies and MMAs serialise in the tensor
    // pipe (measured: ~1000 clk per k-stage with the MMAs removed -- the 12 copies alone -- against 512 clk of MMA)
    static constexpr int kSfSet = kSfaCols + kSfbCols;
    static constexpr int kTmemLd = FP4 ? kTmemSfa + 2 * kSfSet : kNumAcc * BN;  // EPI_QUANT: D2 [128 x Rdp]
    // FP4 mainloops are 4x shorter than INT4 ones, so the epilogue is the bottleneck (measured: the fused
    // fc1 epilogue needs ~5k issue cycles per 128x128 tile against a 3k-cycle mainloop): two half-groups of
    // 4 warps drain the two column halves of a tile concurrently.  INT4 keeps one group (its 8 converter
    // warps already fill the register file).
    static constexpr int kEpiGroups = FP4 ? 2 : 1;
    static constexpr int kEpiThreads = 128 * kEpiGroups;
    static constexpr int kThreads = FP4 ? (4 + 4 * kEpiGroups) * 32 : (kConvWarp0 + kConvWarps) * 32;
    static_assert(!FP4 || kTmemSfa + 2 * kSfSet <= 512, "TMEM budget");
    static_assert(kNumAcc * BN <= 512, "TMEM budget");
    static_assert(EPI != EPI_QUANT || ((BN == 128 || !FP4) && kTmemLd + kMaxRdp <= 512), "fused quantise epilogue: TMEM budget");
    static_assert(EPI != EPI_ROPE || BN == 128, "RMSNorm+RoPE epilogue: one 128-wide head per tile");
    // EPI_LITELA: a 128-wide K | V tile = two heads of [K 32 | V 32]; its [128 x 128] fp32 Gram accumulator sits behind the main accumulators
    static constexpr int kTmemVk = kTmemLd;
    static_assert(EPI != EPI_LITELA || (BN == 128 && kNumAcc == 2 && kTmemVk + 128 <= 512), "LiteLA epilogue: tile shape / TMEM budget");
};

template <bool FP4, int BN, int EPI = EPI_DEFAULT>
struct alignas(1024) Smem {
    using C = Cfg<FP4, BN, EPI>;
    // TMA-staged packed operands
    alignas(1024) uint8_t a[C::kStages][C::kABytes];
    alignas(1024) uint8_t b[C::kStages][C::kBBytes];
    // FP4: scale-factor tiles (tcgen05.cp layout).  INT4: per-group scales (hT)
    alignas(128) uint8_t sa[C::kStages][FP4 ? C::kK64 * 512 : BM * 2];
    alignas(128) uint8_t sb[C::kStages][FP4 ? (BN / 128) * C::kK64 * 512 : BN * 2];
    // INT4: converted hT tiles, 128B-swizzled K-major [rows][64]
    alignas(1024) uint8_t a_cv[FP4 ? 1 : C::kConvStages][FP4 ? 16 : BM * 128];
    alignas(1024) uint8_t b_cv[FP4 ? 1 : C::kConvStages][FP4 ? 16 : BN * 128];
    // low-rank operands, UMMA no-swizzle K-major core-matrix order, one 32-rank chunk
    alignas(1024) uint8_t lora_a[BM * kLoraChunk * 2];
    alignas(1024) uint8_t lora_b[BN * kLoraChunk * 2];
    // epilogue staging for TMA store: [128 rows][64 cols] hT, 128B swizzle, double buffered
    alignas(1024) uint8_t out_stage[2][BM * 128];
    // EPI_QUANT: next layer's lora_down, per 64-column chunk a K-major 128B-swizzled [Rdp][64] tile
    alignas(1024) uint8_t ld_b[EPI == EPI_QUANT ? (BN / 64) * C::kMaxRdp * 128 : 16];
    float bias[BN];
    float cscale[BN];
    float ropesum[2][BM];  // EPI_ROPE with two epilogue groups: per-row partial sums of squares
    float aux[256];  // EPI_QUANT: reciprocals of the next layer's smooth factors of this tile; EPI_ROPE: RMSNorm weights (q | k)
    float aux2[EPI == EPI_QUANT ? 256 : 1];  // EPI_QUANT: 2^24 pre-scale of denormal smooth factors (1 otherwise)
    uint64_t full[C::kStages];
    uint64_t empty[C::kStages];
    uint64_t cfull[C::kConvStages];
    uint64_t cempty[C::kConvStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint64_t lora_b_full;
    uint64_t lora_a_full;
    uint64_t lora_empty;
    uint64_t ld_b_full;
    uint64_t ld_b_empty;
    uint64_t stage_mma_done[2];
    uint64_t d2_full;
    uint64_t vk_full;
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

// ---------------------------------------------------------------------------------------------
// INT4 converter: 16 packed bytes (32 elements) -> 64 bytes of hT in a 128B-swizzled row
// ---------------------------------------------------------------------------------------------
template <typename hT>
__device__ __forceinline__ void convert_unit(uint32_t pk_tile, uint32_t cv_tile, int unit, uint32_t scales, uint32_t offset_bits,
                                             int dbg_nostore = 0) {
    // all operands are SHARED-MEMORY addresses (32-bit): explicit ld.shared / st.shared instead of generic accesses
    using Tr = HalfTraits<hT>;
    using T2 = typename Tr::T2;
    constexpr uint32_t kMagic = Tr::kIsBf16 ? 0x43004300u : 0x64006400u;  // 128 + u  |  1024 + u
    const int r = unit >> 1, h = unit & 1;
    uint32_t words[4];
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(words[0]), "=r"(words[1]), "=r"(words[2]), "=r"(words[3]) : "r"(pk_tile + unit * 16));
    uint16_t sbits;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(sbits) : "r"(scales + r * 2));
    const uint32_t s2bits = static_cast<uint32_t>(sbits) * 0x00010001u;
    const T2 s2 = *reinterpret_cast<const T2 *>(&s2bits);
    const T2 off = *reinterpret_cast<const T2 *>(&offset_bits);
    const uint32_t row = cv_tile + r * 128;
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
        if (dbg_nostore && (o[0] ^ o[1] ^ o[2] ^ o[3]) != 0x12345678u) continue;  // experiment: keep the math, drop the store
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(row + chunk * 16), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool FP4, typename hT, int BN, int EPI>
__global__ void __launch_bounds__(Cfg<FP4, BN, EPI>::kThreads, 1)
gemm_w4a4_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_wgt,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_ld,
                 const GemmParams p) {
    using C = Cfg<FP4, BN, EPI>;
    using S = Smem<FP4, BN, EPI>;
    using Tr = HalfTraits<hT>;
    extern __shared__ uint8_t smem_raw[];
    // align inside the SHARED address space (pointer arithmetic on the __shared__ array): a round trip through uintptr_t
    // makes every access to `s` a generic LD/ST instead of LDS/STS (seen in the ncu source view)
    S &s = *reinterpret_cast<S *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));

    const long long t_kernel0 = clock64();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int k64_total = p.K >> 6;
    const int num_kblocks = (k64_total + C::kK64 - 1) / C::kK64;
    const int lora_chunks = p.lora_up != nullptr ? p.Rp / kLoraChunk : 0;
    // tile schedule: round-robin (tile = cta + i * grid) or, for the fused down projection, one contiguous range per
    // CTA so that consecutive tiles share the m-block and the projection keeps accumulating in TMEM
    const int tile_begin = p.tile_contig ? static_cast<int>(static_cast<long long>(p.num_tiles) * blockIdx.x / gridDim.x) : blockIdx.x;
    const int tile_end = p.tile_contig ? static_cast<int>(static_cast<long long>(p.num_tiles) * (blockIdx.x + 1) / gridDim.x) : p.num_tiles;
    const int tile_step = p.tile_contig ? 1 : gridDim.x;

    // ---- one-time setup -----------------------------------------------------------------------
    if (warp == 0 && elect_one()) {
        prefetch_tensormap(&tm_act);
        prefetch_tensormap(&tm_wgt);
        prefetch_tensormap(&tm_out);
        if constexpr (EPI == EPI_QUANT) prefetch_tensormap(&tm_ld);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < C::kStages; i++) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], FP4 ? 1 : kNumConvThreads / 32);
        }
        for (int i = 0; i < C::kConvStages; i++) {
            mbar_init(&s.cfull[i], kNumConvThreads / 32);
            mbar_init(&s.cempty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&s.tmem_full[i], 1);
            mbar_init(&s.tmem_empty[i], C::kEpiThreads);
        }
        mbar_init(&s.lora_b_full, 1);
        mbar_init(&s.lora_a_full, C::kEpiThreads);
        mbar_init(&s.lora_empty, 1);
        mbar_init(&s.ld_b_full, 1);
        mbar_init(&s.ld_b_empty, C::kEpiGroups);
        mbar_init(&s.stage_mma_done[0], 1);
        mbar_init(&s.stage_mma_done[1], 1);
        mbar_init(&s.d2_full, C::kEpiGroups);
        mbar_init(&s.vk_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(&s.tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = s.tmem_base;
    const long long t_setup = clock64() - t_kernel0;
    griddep_launch_dependents();   // the next kernel in the stream may start its own setup on SMs this grid leaves idle
    griddep_wait();                // ... and this one touches global memory only after its predecessor has completed

    // INT4: 1024 threads start with 64 registers each; every warp group (4 warps) re-sizes its share as the first
    // instruction of its role: producer/MMA/alloc group 56, epilogue group 128, the six converter groups 48.
    if (warp < 4) {
      if constexpr (!FP4) setmaxnreg_dec<56>();
      if (warp == 0) {
        // =================================== TMA producer =======================================
        if (elect_one()) {
            PipeState st;
            uint32_t lora_phase = 0;
            uint32_t ld_phase = 0;
            long long t_empty = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
                const int mb = tile / p.num_n_blocks, nb = tile % p.num_n_blocks;
                const int m0 = mb * BM, n0 = nb * BN;
                for (int kb = 0; kb < num_kblocks; kb++) {
                    NB200_TIMED(t_empty, mbar_wait(&s.empty[st.idx], st.phase ^ 1));
                    if constexpr (FP4) {
                        const int nj = min(C::kK64, k64_total - C::kK64 * kb);
                        const uint32_t sf_bytes = nj * 512;
                        mbar_expect_tx(&s.full[st.idx], C::kABytes + C::kBBytes + sf_bytes * (1 + BN / 128));
                        tma_load_2d(s.a[st.idx], &tm_act, &s.full[st.idx], kb * 32 * C::kK64, m0);
                        tma_load_2d(s.b[st.idx], &tm_wgt, &s.full[st.idx], kb * 32 * C::kK64, n0);
                        bulk_load(s.sa[st.idx], p.sfa + (static_cast<size_t>(mb) * k64_total + C::kK64 * kb) * 512, sf_bytes,
                                  &s.full[st.idx]);
#pragma unroll
                        for (int h = 0; h < BN / 128; h++)
                            bulk_load(s.sb[st.idx] + h * C::kK64 * 512,
                                      p.sfb + (static_cast<size_t>(n0 / 128 + h) * k64_total + C::kK64 * kb) * 512, sf_bytes,
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
                if constexpr (EPI == EPI_QUANT) {
                    // next layer's lora_down slab: needed only by this tile's EPILOGUE, and its buffer is released by
                    // the previous tile's epilogue -> fetch it after the main-loop stages are in flight, otherwise the
                    // main loop of tile i+1 cannot start before the epilogue of tile i has finished
                    if (p.R_down > 0) {
                        mbar_wait(&s.ld_b_empty, ld_phase ^ 1);
                        mbar_expect_tx(&s.ld_b_full, (BN / 64) * p.Rdp * 128);
#pragma unroll
                        for (int ch = 0; ch < BN / 64; ch++)
                            tma_load_2d(s.ld_b + ch * C::kMaxRdp * 128, &tm_ld, &s.ld_b_full, n0 + ch * 64, 0);
                        ld_phase ^= 1;
                    }
                }
            }
            if (p.prof) p.prof[blockIdx.x * 16 + 0] = t_empty;
        }
      } else if (warp == 1) {
        // ==================================== MMA issuer ========================================
        if (elect_one()) {
            PipeState st;   // FP4: TMA ring.  INT4: converted ring
            uint32_t lora_phase = 0;
            uint32_t acc_phase[2] = {0, 0};
            int it = 0;
            long long t_tmem_empty = 0, t_full = 0, t_lora = 0, t_first = 0;
            const long long t_mma0 = clock64();
            constexpr uint32_t idesc_main = FP4 ? make_idesc_nvf4(BM, BN) : make_idesc_f16(Tr::kIsBf16, BM, BN);
            constexpr uint32_t idesc_lora = make_idesc_f16(Tr::kIsBf16, BM, BN);
            [[maybe_unused]] const bool lean = p.prof == nullptr && p.debug == 0;
            [[maybe_unused]] const uint64_t adesc0 = make_sw128_kmajor_desc(smem_u32(s.a[0])), bdesc0 = make_sw128_kmajor_desc(smem_u32(s.b[0]));
            [[maybe_unused]] const uint64_t sadesc0 = make_smem_desc(smem_u32(s.sa[0]), 0, 128, kLayoutNoSwizzle);
            [[maybe_unused]] const uint64_t sbdesc0 = make_smem_desc(smem_u32(s.sb[0]), 0, 128, kLayoutNoSwizzle);
            for (int tile = tile_begin; tile < tile_end; tile += tile_step, it++) {
                const int acc = it % C::kNumAcc;
                NB200_TIMED(t_tmem_empty, mbar_wait(&s.tmem_empty[acc], acc_phase[acc] ^ 1));
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kblocks; kb++) {
                    if constexpr (FP4 && C::kK64 == 4) {
                        // Hot path (no profiling counters, no ablation bits, full k block): this ONE thread feeds the tensor pipe, so
                        // descriptors are base + slot * stride, the copies and MMAs are unrolled and there are no clock reads (r02:
                        // the timeline of the cluster kernel showed 618 clk to issue a stage against 512 clk of MMA time).
                        if (lean && k64_total - C::kK64 * kb >= 4) {
                            mbar_wait(&s.full[st.idx], st.phase);
                            tc_fence_after_sync();
                            const uint32_t sf_set = tmem_base + (kb & 1) * C::kSfSet;
                            const uint64_t ad = adesc0 + static_cast<uint64_t>(st.idx) * (C::kABytes >> 4), bd = bdesc0 + static_cast<uint64_t>(st.idx) * (C::kBBytes >> 4);
                            const uint64_t sad = sadesc0 + static_cast<uint64_t>(st.idx) * ((C::kK64 * 512) >> 4);
                            const uint64_t sbd = sbdesc0 + static_cast<uint64_t>(st.idx) * (((BN / 128) * C::kK64 * 512) >> 4);
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                tc_cp_32x128b_warpx4(sf_set + C::kTmemSfa + 4 * j, sad + 32 * j);
#pragma unroll
                                for (int h = 0; h < BN / 128; h++)
                                    tc_cp_32x128b_warpx4(sf_set + C::kTmemSfb + (BN / 32) * j + 4 * h, sbd + h * ((C::kK64 * 512) >> 4) + 32 * j);
                            }
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                tc_mma_nvf4(tmem_d, ad + 2 * j, bd + 2 * j, idesc_main, sf_set + C::kTmemSfa + 4 * j, sf_set + C::kTmemSfb + (BN / 32) * j, (kb | j) != 0);
                            tc_commit(&s.empty[st.idx]);
                            st.advance(C::kStages);
                            continue;
                        }
                    }
                    if constexpr (FP4) {
                        NB200_TIMED(t_full, mbar_wait(&s.full[st.idx], st.phase));
                        if (t_first == 0) t_first = clock64() - t_mma0;
                        tc_fence_after_sync();
                        const int nj = min(C::kK64, k64_total - C::kK64 * kb);
                        const uint32_t sf_set = tmem_base + (kb & 1) * C::kSfSet;
                        // operand rows are 32 * kK64 bytes: 128-byte rows -> SWIZZLE_128B atoms of 1024 B, 64-byte rows -> SWIZZLE_64B
                        // atoms of 512 B; the K64 block j starts 32 * j bytes into the row
                        auto op_desc = [](uint32_t addr) {
                            return C::kK64 == 4 ? make_sw128_kmajor_desc(addr) : make_smem_desc(addr, 16, 512, kLayoutSw64);
                        };
                        for (int j = 0; j < nj; j++) {
                            tc_cp_32x128b_warpx4(sf_set + C::kTmemSfa + 4 * j,
                                                 make_smem_desc(smem_u32(s.sa[st.idx] + j * 512), 0, 128, kLayoutNoSwizzle));
#pragma unroll
                            for (int h = 0; h < BN / 128; h++)
                                tc_cp_32x128b_warpx4(
                                    sf_set + C::kTmemSfb + (BN / 32) * j + 4 * h,
                                    make_smem_desc(smem_u32(s.sb[st.idx] + h * C::kK64 * 512 + j * 512), 0, 128, kLayoutNoSwizzle));
                        }
                        const uint32_t a_addr = smem_u32(s.a[st.idx]), b_addr = smem_u32(s.b[st.idx]);
                        for (int j = 0; j < ((p.debug & 4) ? 0 : nj); j++)
                            tc_mma_nvf4(tmem_d, op_desc(a_addr + j * 32), op_desc(b_addr + j * 32), idesc_main, sf_set + C::kTmemSfa + 4 * j, sf_set + C::kTmemSfb + (BN / 32) * j,
                                        (kb | j) != 0);
                        tc_commit(&s.empty[st.idx]);
                        st.advance(C::kStages);
                    } else {
                        NB200_TIMED(t_full, mbar_wait(&s.cfull[st.idx], st.phase));
                        if (t_first == 0) t_first = clock64() - t_mma0;
                        tc_fence_after_sync();
                        const uint32_t a_addr = smem_u32(s.a_cv[st.idx]), b_addr = smem_u32(s.b_cv[st.idx]);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (!(p.debug & 4))
                                tc_mma_f16(tmem_d, make_sw128_kmajor_desc(a_addr + j * 32), make_sw128_kmajor_desc(b_addr + j * 32),
                                           idesc_main, (kb | j) != 0);
                        tc_commit(&s.cempty[st.idx]);
                        st.advance(C::kConvStages);
                    }
                }
                for (int c = 0; c < lora_chunks; c++) {
                    NB200_TIMED(t_lora, mbar_wait(&s.lora_b_full, lora_phase));
                    NB200_TIMED(t_lora, mbar_wait(&s.lora_a_full, lora_phase));
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
            if (p.prof) {
                p.prof[blockIdx.x * 16 + 1] = t_tmem_empty;
                p.prof[blockIdx.x * 16 + 2] = t_full;
                p.prof[blockIdx.x * 16 + 3] = t_lora;
                p.prof[blockIdx.x * 16 + 10] = clock64() - t_mma0;
                p.prof[blockIdx.x * 16 + 13] = t_first;
            }
        }
      }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4 * C::kEpiGroups) {
        if constexpr (!FP4) setmaxnreg_inc<128>();
        // ===================================== epilogue ==========================================
        constexpr int H = C::kEpiGroups;       // half-groups of 4 warps; group h owns column chunks [h*CH/H, (h+1)*CH/H)
        constexpr int CH = BN / 64;
        const int h = (warp - kEpiWarp0) >> 2;
        const int q = warp & 3;                // TMEM lane quadrant (hardware: warp id % 4)
        const int et = threadIdx.x - (kEpiWarp0 + 4 * h) * 32;   // thread within the group
        const int eta = threadIdx.x - kEpiWarp0 * 32;            // thread within all epilogue warps
        const int row = q * 32 + lane;         // row inside the tile
        uint32_t lora_phase = 0;
        uint32_t acc_phase[2] = {0, 0};
        int it = 0;
        uint32_t store_count = 0;
        long long t_tmem_full = 0, t_pre = 0;
        const long long t_epi0 = clock64();
        [[maybe_unused]] uint32_t ld_phase = 0, d2_phase = 0, vk_phase = 0;
        [[maybe_unused]] uint32_t smd_phase[2] = {0, 0};
        if constexpr (EPI == EPI_ROPE) {
            for (int i = eta; i < 256; i += C::kEpiThreads)
                s.aux[i] = Tr::to_float(reinterpret_cast<const hT *>(i < 128 ? p.norm_q : p.norm_k)[i & 127]);
        }
        // Software pipelining of the per-tile global loads (channel vectors, first chunk of the low-rank activations):
        // they are issued one tile ahead and sit in registers while the current tile's chunks are processed -- two dependent
        // L2 round trips (~3k clk per tile, a quarter of the fused epilogue) no longer sit on the epilogue's serial path.
        constexpr int kVecIters = (BN + C::kEpiThreads - 1) / C::kEpiThreads;
        constexpr bool kPrefetchLora = EPI == EPI_QUANT && FP4;   // register budget: only where the epilogue is the bottleneck
        float pf_bias[kVecIters], pf_cs[kVecIters];
        [[maybe_unused]] float pf_sm[kVecIters];
        [[maybe_unused]] float4 pf_lora[kPrefetchLora ? 2 * (4 / H) : 1];
        auto prefetch = [&](int t) {
            if (t >= tile_end) return;
            const int n0_ = (t % p.num_n_blocks) * BN, m0_ = (t / p.num_n_blocks) * BM;
#pragma unroll
            for (int k = 0; k < kVecIters; k++) {
                const int i = eta + k * C::kEpiThreads;
                if (i < BN) {
                    pf_bias[k] = p.bias != nullptr ? p.bias[n0_ + i] : 0.f;
                    pf_cs[k] = p.cscale != nullptr ? p.cscale[n0_ + i] : 1.f;
                    if constexpr (EPI == EPI_QUANT) pf_sm[k] = Tr::to_float(reinterpret_cast<const hT *>(p.smooth_next)[n0_ + i]);
                }
            }
            if constexpr (kPrefetchLora) {
                if (lora_chunks > 0) {
                    const float *src = p.lora_act + static_cast<size_t>(m0_ + row) * p.R;
#pragma unroll
                    for (int oo = 0; oo < 4 / H; oo++) {
                        const int o = h * (4 / H) + oo;
                        if (o * 8 < p.R) {
                            pf_lora[2 * oo] = *reinterpret_cast<const float4 *>(src + o * 8);
                            pf_lora[2 * oo + 1] = *reinterpret_cast<const float4 *>(src + o * 8 + 4);
                        }
                    }
                }
            }
        };
        prefetch(tile_begin);
        for (int tile = tile_begin; tile < tile_end; tile += tile_step, it++) {
            const int mb = tile / p.num_n_blocks, nb = tile % p.num_n_blocks;
            const int m0 = mb * BM, n0 = nb * BN;
            const int acc = it % C::kNumAcc;
            const long long t_tile0 = clock64();

            // per-tile channel vectors (prefetched one tile ahead)
            named_bar_sync(1, C::kEpiThreads);
#pragma unroll
            for (int k = 0; k < kVecIters; k++) {
                const int i = eta + k * C::kEpiThreads;
                if (i < BN) {
                    s.bias[i] = pf_bias[k];
                    s.cscale[i] = pf_cs[k];
                    if constexpr (EPI == EPI_QUANT) {
                        // x / smooth is the reference's __fdividef (gemm_w4a4.cuh:930-1043): SASS "if |b| < 2^-126 scale a and b by
                        // 2^24; MUFU.RCP(b) * a".  Reciprocal and pre-scale depend only on the column: once per tile, not per row.
                        const float b = pf_sm[k];
                        const bool tiny = fabsf(b) < 1.175494350822287508e-38f;
                        s.aux2[i] = tiny ? 16777216.f : 1.f;
                        s.aux[i] = rcp_approx(tiny ? b * 16777216.f : b);
                    }
                }
            }
            named_bar_sync(1, C::kEpiThreads);

            // low-rank activations: fp32 -> * lora_scale -> hT  (lora.cuh:145-151), one row per thread
            for (int c = 0; c < lora_chunks; c++) {
                mbar_wait(&s.lora_empty, lora_phase ^ 1);
                const float *src = p.lora_act + static_cast<size_t>(m0 + row) * p.R + c * kLoraChunk;
                uint8_t *dst = s.lora_a + (row >> 3) * 512 + (row & 7) * 16;
#pragma unroll
                for (int oo = 0; oo < 4 / H; oo++) {  // rank octets of this group
                    const int o = h * (4 / H) + oo;
                    const int r0 = c * kLoraChunk + o * 8;
                    uint32_t w[4] = {0, 0, 0, 0};
                    if (r0 < p.R) {
                        float4 f0, f1;
                        if (kPrefetchLora && c == 0) {
                            f0 = pf_lora[2 * oo];
                            f1 = pf_lora[2 * oo + 1];
                        } else {
                            f0 = *reinterpret_cast<const float4 *>(src + o * 8);
                            f1 = *reinterpret_cast<const float4 *>(src + o * 8 + 4);
                        }
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
            prefetch(tile + tile_step);   // in flight while this tile's chunks are processed

            t_pre += clock64() - t_tile0;
            NB200_TIMED(t_tmem_full, mbar_wait(&s.tmem_full[acc], acc_phase[acc]));
            acc_phase[acc] ^= 1;
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

            // ---- EPI_ROPE pass 1: per-row sum of squares over the 128-wide head (epilogues.cuh:327-341)
            // fused down projection: D2 keeps accumulating in TMEM while consecutive tiles stay in the same m-block
            [[maybe_unused]] const bool d2_fresh = tile == tile_begin || (tile - tile_step) / p.num_n_blocks != mb;
            [[maybe_unused]] const bool d2_flush = tile + tile_step >= tile_end || (tile + tile_step) / p.num_n_blocks != mb;
            [[maybe_unused]] bool do_rope = false;
            [[maybe_unused]] int qkv_part = 0, qkv_head = 0;
            [[maybe_unused]] float rope_coef = 1.f;
            [[maybe_unused]] const float *normw = s.aux;
            [[maybe_unused]] const float *rot_row = nullptr;
            [[maybe_unused]] bool vk_tile = false;   // EPI_LITELA: a K | V tile (reduced into out_vk, nothing stored) vs a Q tile (relu, stored)
            if constexpr (EPI == EPI_LITELA) vk_tile = nb >= p.num_n_blocks / 3;
            if constexpr (EPI == EPI_ROPE) {
                const int part = nb / (p.num_n_blocks / 3);  // 0 = Q heads, 1 = K heads, 2 = V (untouched)
                qkv_part = part;
                qkv_head = nb % (p.num_n_blocks / 3);
                do_rope = part < 2;
                if (do_rope) {
                    // 8 partial sums of 16 columns (sequential FMA chains) combined by the fixed tree ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)): the
                    // order of csrc/rope.cu's 8-lanes-per-head reduction and of the cluster kernel's epilogue (bit-identical routes)
                    float sumsq = 0.f, pair = 0.f;
#pragma unroll 1
                    for (int cc = 0; cc < 4 / H; cc++) {
                        const int c32 = h * (4 / H) + cc;
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + c32 * 32, v);
                        tmem_ld_wait();
                        float pa = 0.f, pb = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float y0 = fmaf(__uint_as_float(v[i]), s.cscale[c32 * 32 + i], s.bias[c32 * 32 + i]);
                            const float y1 = fmaf(__uint_as_float(v[i + 1]), s.cscale[c32 * 32 + i + 1], s.bias[c32 * 32 + i + 1]);
                            const float2 r = Tr::to_float2(Tr::from_float2(make_float2(y0, y1)));  // fpsum is hT
                            if (i < 16) {
                                pa = fmaf(r.x, r.x, pa);
                                pa = fmaf(r.y, r.y, pa);
                            } else {
                                pb = fmaf(r.x, r.x, pb);
                                pb = fmaf(r.y, r.y, pb);
                            }
                        }
                        if ((cc & 1) == 0) {
                            pair = pa + pb;
                        } else {
                            const float quad = pair + (pa + pb);
                            sumsq = cc == 1 ? quad : sumsq + quad;
                        }
                    }
                    if constexpr (H == 2) {  // each group saw 64 of the head's 128 columns
                        s.ropesum[h][row] = sumsq;
                        named_bar_sync(1, C::kEpiThreads);
                        sumsq = s.ropesum[0][row] + s.ropesum[1][row];
                    }
                    rope_coef = rsqrt_approx_ftz(sumsq / 128.f + 1e-6f);
                    normw = s.aux + part * 128;
                    // reference pack_rotemb order (transformer_flux.py:60-92): float index of (row m, pair pr, sin|cos)
                    //   ((((m/16*16 + pr/4)*8 + m%8)*4 + pr%4)*2 + (m%16)/8)*2 + {0,1}
                    const int m = m0 + row;
                    rot_row = p.rotary + (static_cast<size_t>(m >> 4) * 16 * 8 + (m & 7)) * 16 + ((m >> 3) & 1) * 2;
                }
            }
            if constexpr (EPI == EPI_QUANT) {
                if (p.R_down > 0) {
                    mbar_wait(&s.ld_b_full, ld_phase);
                    ld_phase ^= 1;
                }
            }

            // one 64-column chunk: scale/bias/activation, hT pack into the swizzled staging tile, TMA store
            auto do_chunk = [&](const int ch, const uint32_t (&v0)[32], const uint32_t (&v1)[32]) {
                // staging buffer: one group -> two buffers alternate; two groups -> one buffer each
                const int buf = H == 2 ? h : (store_count & 1);
                [[maybe_unused]] bool pack_qkv = false;
                if constexpr (EPI == EPI_ROPE) pack_qkv = p.out_qkv[0] != nullptr;
                if (et == 0 && !pack_qkv) {
                    if constexpr (H == 2) {
                        bulk_wait_group_read<0>();
                    } else {
                        if (vk_tile)   // (a K | V tile commits no store group: "all but the newest" would not cover the other buffer's store)
                            bulk_wait_group_read<0>();
                        else
                            bulk_wait_group_read<1>();
                    }
                }
                if constexpr (EPI == EPI_QUANT) {
                    if (p.R_down > 0) mbar_wait(&s.stage_mma_done[buf], smd_phase[buf] ^ 1);
                }
                if (!pack_qkv) named_bar_sync(2 + 2 * h, kNumEpiThreads);
                uint8_t *srow = s.out_stage[buf] + row * 128;
                // EpiloguePackQKV (epilogues.cuh:427-550): this thread's row of the head, 64 fp16 of this chunk
                [[maybe_unused]] __half *qkv_row = nullptr;
                [[maybe_unused]] bool qkv_masked = false;
                if constexpr (EPI == EPI_ROPE) {
                    if (pack_qkv) {
                        qkv_row = p.out_qkv[qkv_part] + qkv_head * p.stride_head[qkv_part] + static_cast<size_t>(m0 + row) * 128 + ch * 64;
                        qkv_masked = m0 + row >= p.attn_tokens;
                    }
                }
                [[maybe_unused]] uint32_t gw[32];  // EPI_QUANT: the 64 hT values of this row/chunk
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
                        if constexpr (EPI == EPI_QUANT) {
                            // fused mode always applies GELU (launch_impl:296-308) to the hT-rounded value
                            const float2 r = Tr::to_float2(Tr::from_float2(make_float2(y0, y1)));
                            y0 = gelu_f32(r.x);
                            y1 = gelu_f32(r.y);
                        } else if constexpr (EPI == EPI_ROPE) {
                            if (do_rope) {
                                const float2 r = Tr::to_float2(Tr::from_float2(make_float2(y0, y1)));
                                const int hc = ch * 64 + col;  // column inside the head; pair index hc / 2
                                const float x0 = r.x * (rope_coef * normw[hc]);
                                const float x1 = r.y * (rope_coef * normw[hc + 1]);
                                const int pr = hc >> 1;
                                const float2 sc = *reinterpret_cast<const float2 *>(rot_row + (pr >> 2) * 128 + (pr & 3) * 4);
                                y0 = x0 * sc.y - x1 * sc.x;  // (sin, cos) = (sc.x, sc.y)  epilogues.cuh:362-367
                                y1 = x0 * sc.x + x1 * sc.y;
                            }
                        } else if constexpr (EPI == EPI_LITELA) {
                            // relu on the hT value: every column of a Q tile (epilogues.cuh:676-688), the K half (first 32 columns) of each
                            // head of a K | V tile (epilogues.cuh:611-613)
                            if (!vk_tile || col < 32) {
                                const float2 r = Tr::to_float2(Tr::from_float2(make_float2(y0, y1)));
                                y0 = fmaxf(r.x, 0.f);
                                y1 = fmaxf(r.y, 0.f);
                            }
                        } else {
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
                        }
                        if constexpr (!Tr::kIsBf16 && EPI != EPI_QUANT) {  // fp16 stores clamp (gemm_base.cuh:688-696)
                            y0 = fminf(fmaxf(y0, -65504.f), 65504.f);
                            y1 = fminf(fmaxf(y1, -65504.f), 65504.f);
                        }
                        typename Tr::T2 h = Tr::from_float2(make_float2(y0, y1));
                        w[i] = *reinterpret_cast<uint32_t *>(&h);
                        if constexpr (EPI == EPI_QUANT) gw[c8 * 4 + i] = w[i];
                        if constexpr (EPI == EPI_ROPE) {
                            if (pack_qkv) {  // hT -> fp16 through fp32 (convert_half2, epilogues.cuh:446-453); pad rows masked
                                const __half2 hh = __float22half2_rn(Tr::to_float2(h));
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
                if constexpr (EPI == EPI_ROPE) {
                    if (pack_qkv) {
                        store_count++;
                        return;
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(3 + 2 * h, kNumEpiThreads);
                if constexpr (EPI == EPI_QUANT && H == 2) {
                    // the second group's down-projection MMAs accumulate onto the first group's (accumulate = 0):
                    // issue them in order.  Named barrier 6: group 0's issuing warp arrives after issuing.
                    if (p.R_down > 0 && h == 1 && q == 0) asm volatile("bar.sync 6, 64;" ::: "memory");
                }
                if (et == 0) {
                    if ((EPI != EPI_QUANT || p.out != nullptr) && !vk_tile) {
                        tma_store_2d(&tm_out, s.out_stage[buf], n0 + ch * 64, m0);
                        bulk_commit_group();
                    }
                    if constexpr (EPI == EPI_QUANT) {
                        if (p.R_down > 0) {
                            // next layer's low-rank down projection on the GELU output (no shift), lora.cuh:243-353:
                            // D2[128 x Rdp] += G[128 x 64] * Ld_next[Rdp x 64]^T, staged tile doubles as the A operand
                            tc_fence_after_sync();
                            const uint32_t a_addr = smem_u32(s.out_stage[buf]);
                            const uint32_t b_addr = smem_u32(s.ld_b + ch * C::kMaxRdp * 128);
                            const uint32_t idesc_ld = make_idesc_f16(Tr::kIsBf16, BM, p.Rdp);
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                tc_mma_f16(tmem_base + C::kTmemLd, make_sw128_kmajor_desc(a_addr + j * 32),
                                           make_sw128_kmajor_desc(b_addr + j * 32), idesc_ld, (ch | j) != 0 || !d2_fresh);
                            tc_commit(&s.stage_mma_done[buf]);
                        }
                    }
                }
                if constexpr (EPI == EPI_QUANT && H == 2) {
                    if (p.R_down > 0 && h == 0 && q == 0) asm volatile("bar.arrive 6, 64;" ::: "memory");
                }
                if constexpr (EPI == EPI_QUANT) {
                    smd_phase[buf] ^= 1;
                    if (p.debug & 32) {
                        store_count++;
                        return;
                    }
                    // ---- next-layer quantisation of (gelu + shift) / smooth  (gemm_w4a4.cuh:930-1043) ---------
                    const int m = m0 + row;
                    uint8_t *qrow = p.qout + static_cast<size_t>(m) * (p.N >> 1) + ((n0 + ch * 64) >> 1);
                    const float *smo = s.aux + ch * 64;    // rcp.approx of the smooth factors
                    const float *smk = s.aux2 + ch * 64;   // their denormal pre-scale
                    if constexpr (!FP4) {
                        typename Tr::T2 sh2;
                        sh2.x = Tr::from_float(0.171875f);
                        sh2.y = sh2.x;
                        float amax = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const typename Tr::T2 gsh = __hadd2(*reinterpret_cast<typename Tr::T2 *>(&gw[i]), sh2);
                            const float2 f = Tr::to_float2(gsh);
                            const typename Tr::T2 d = Tr::from_float2(
                                make_float2((f.x * smk[2 * i]) * smo[2 * i], (f.y * smk[2 * i + 1]) * smo[2 * i + 1]));
                            gw[i] = *reinterpret_cast<const uint32_t *>(&d);
                            const float2 df = Tr::to_float2(d);
                            amax = fmaxf(amax, fmaxf(fabsf(df.x), fabsf(df.y)));
                        }
                        const float s32 = amax * (1.0f / 15.0f);
                        const float rs = rcp_approx_ftz(s32);
                        reinterpret_cast<hT *>(p.oscales_out)[static_cast<size_t>((n0 >> 6) + ch) * p.Mp + m] = Tr::from_float(s32);
                        uint32_t words[8];
#pragma unroll
                        for (int w8 = 0; w8 < 8; w8++) {
                            int qv[8];
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const float2 df = Tr::to_float2(*reinterpret_cast<typename Tr::T2 *>(&gw[w8 * 4 + i]));
                                qv[2 * i] = cvt_rni(df.x * rs);
                                qv[2 * i + 1] = cvt_rni(df.y * rs);
                            }
                            words[w8] = pack8_int4_b200<true>(qv);
                        }
                        *reinterpret_cast<uint4 *>(qrow) = make_uint4(words[0], words[1], words[2], words[3]);
                        *reinterpret_cast<uint4 *>(qrow + 16) = make_uint4(words[4], words[5], words[6], words[7]);
                    } else {
                        uint32_t words[8];
                        uint32_t sfw = 0;
#pragma unroll
                        for (int g16 = 0; g16 < 4; g16++) {
                            float amax = 0.f;
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const int wi = g16 * 8 + i;
                                const float2 f = Tr::to_float2(*reinterpret_cast<typename Tr::T2 *>(&gw[wi]));  // shift is 0 for FP4
                                const typename Tr::T2 d = Tr::from_float2(
                                    make_float2((f.x * smk[2 * wi]) * smo[2 * wi], (f.y * smk[2 * wi + 1]) * smo[2 * wi + 1]));
                                gw[wi] = *reinterpret_cast<const uint32_t *>(&d);
                                const float2 df = Tr::to_float2(d);
                                amax = fmaxf(amax, fmaxf(fabsf(df.x), fabsf(df.y)));
                            }
                            const float sc = fminf(amax * (1.0f / 6.0f), 448.0f);
                            const float rs = rcp_approx_ftz(sc);
                            sfw |= (cvt_e4m3x2(0.f, sc) & 0xFFu) << (8 * g16);
#pragma unroll
                            for (int w2 = 0; w2 < 2; w2++) {
                                uint32_t wv = 0;
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    const float2 df = Tr::to_float2(*reinterpret_cast<typename Tr::T2 *>(&gw[g16 * 8 + w2 * 4 + i]));
                                    wv |= cvt_e2m1x2(df.y * rs, df.x * rs) << (8 * i);
                                }
                                words[g16 * 2 + w2] = wv;
                            }
                        }
                        *reinterpret_cast<uint4 *>(qrow) = make_uint4(words[0], words[1], words[2], words[3]);
                        *reinterpret_cast<uint4 *>(qrow + 16) = make_uint4(words[4], words[5], words[6], words[7]);
                        uint8_t *sf = reinterpret_cast<uint8_t *>(p.oscales_out) +
                                      (static_cast<size_t>(m >> 7) * (p.N >> 6) + (n0 >> 6) + ch) * 512 + (m & 31) * 16 +
                                      ((m & 127) >> 5) * 4;
                        *reinterpret_cast<uint32_t *>(sf) = sfw;
                    }
                }
                store_count++;
            };
            // FP4 with a single 256-column accumulator: pull this group's two chunks into registers first and
            // hand the accumulator back to the MMA warp before doing any math (the exposed part of the epilogue
            // shrinks from "drain + math + stores" to four tcgen05.ld).
            constexpr bool kEarlyRelease = FP4 && C::kNumAcc == 1 && CH / H == 2 && EPI == EPI_DEFAULT;
            if constexpr (kEarlyRelease) {
                uint32_t va[32], vb[32], vc[32], vd[32];
                const int ch0 = h * 2;
                tmem_ld_32x32b_x32(taddr + ch0 * 64, va);
                tmem_ld_32x32b_x32(taddr + ch0 * 64 + 32, vb);
                tmem_ld_32x32b_x32(taddr + ch0 * 64 + 64, vc);
                tmem_ld_32x32b_x32(taddr + ch0 * 64 + 96, vd);
                tmem_ld_wait();
                tc_fence_before_sync();
                mbar_arrive(&s.tmem_empty[acc]);
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
                        mbar_arrive(&s.tmem_empty[acc]);
                    }
                    if (!(p.debug & 8)) do_chunk(ch, v0, v1);
                }
            }
            if constexpr (EPI == EPI_LITELA) {
                if (vk_tile) {
                    // EpilogueLiteLA (epilogues.cuh:571-660): the tile's two heads, [K 32 | V 32] each, are staged as hT [128 tokens][64] per head in the
                    // two store buffers (relu already on K).  The token reduction  vk[v][k] = sum_t V[t,v] relu(K[t,k])  is a GEMM with the TOKENS as
                    // its K dimension: both operands are the staged tile read as an MN-major SW128 operand (the layout csrc/attention.cu feeds V
                    // with), D[128 x 128] = X^T X in fp32 -- the reference's mma_f16xf16_f32 on transposed fragments.  Head hd's state is the
                    // block D[64 hd + 32 .. 64 hd + 63][64 hd .. 64 hd + 31] = TMEM lanes of the warps q = 1 / q = 3, 32 columns each; the
                    // other blocks (K^T K, V^T V, cross-head) are computed and ignored: 8 instructions of 64 clk per tile.
                    named_bar_sync(1, C::kEpiThreads);   // both heads staged (every thread fenced its writes for the async proxy before its group barrier)
                    if (eta == 0) {
                        tc_fence_after_sync();
                        const uint32_t x_addr = smem_u32(s.out_stage[0]);
                        constexpr uint32_t idesc_vk = make_idesc_f16(Tr::kIsBf16, 128, 128) | (1u << 15) | (1u << 16);   // A and B MN-major
#pragma unroll
                        for (int ks = 0; ks < BM / 16; ks++) {   // 16 tokens per instruction: 2 groups of 8 rows = 2 KB further
                            const uint64_t xd = make_smem_desc(x_addr + ks * 2048, BM * 128, 1024, kLayoutSw128);   // LBO = distance between the two heads' buffers
                            tc_mma_f16(tmem_base + C::kTmemVk, xd, xd, idesc_vk, ks != 0);
                        }
                        tc_commit(&s.vk_full);
                    }
                    const int batch = m0 / p.vk_tokens;
                    float *vk_tile_out = p.out_vk + (static_cast<size_t>(batch) * p.vk_heads + (nb - p.num_n_blocks / 3) * 2) * (33 * 32);
                    // row 32 of the state, sum_t relu(K[t, k]) (the reference multiplies by a fragment of ones, epilogues.cuh:626-655): column sums
                    // of the staged K halves on the CUDA cores while the tensor core works -- thread = (column j, quarter of the rows)
                    {
                        const int j = et & 31, part = et >> 5;
#pragma unroll 1
                        for (int cc = 0; cc < CH / H; cc++) {
                            const int hd = h * (CH / H) + cc;
                            const uint8_t *xt = s.out_stage[hd];
                            float sum = 0.f;
#pragma unroll 8
                            for (int r = 0; r < 32; r++) {
                                const int t = part * 32 + r;
                                sum += Tr::to_float(*reinterpret_cast<const hT *>(xt + t * 128 + (((j >> 3) ^ (t & 7)) << 4) + (j & 7) * 2));
                            }
                            atomicAdd(vk_tile_out + hd * (33 * 32) + 32 * 32 + j, sum);
                        }
                    }
                    mbar_wait(&s.vk_full, vk_phase);
                    vk_phase ^= 1;
                    tc_fence_after_sync();
                    if (h == 0 && (q & 1)) {
                        const int hd = q >> 1;
                        uint32_t d[32];
                        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + C::kTmemVk + hd * 64, d);
                        tmem_ld_wait();
                        float *dst = vk_tile_out + hd * (33 * 32) + lane * 32;   // this thread's V index = lane
#pragma unroll
                        for (int i = 0; i < 32; i++) atomicAdd(dst + i, __uint_as_float(d[i]));
                    }
                    tc_fence_before_sync();   // (the next tile's barrier orders these TMEM / shared-memory reads before the buffers are written again)
                }
            }
            if constexpr (EPI == EPI_QUANT) {
                if (p.R_down > 0) {
                    if (et == 0) {
                        if (d2_flush) tc_commit(&s.d2_full);
                        tc_commit(&s.ld_b_empty);
                    }
                    if (d2_flush) {  // last tile of this CTA's run in the m-block: add the partial projection to HBM
                        mbar_wait(&s.d2_full, d2_phase);
                        d2_phase ^= 1;
                        tc_fence_after_sync();
                        float *dst = p.lora_act_out + static_cast<size_t>(m0 + row) * p.R_down;
                        if (p.ws_partial != nullptr) {
                            // Deterministic reduction (the reference adds its partials with red.global.add.f32, launch_impl:252 +
                            // lora.cuh:320-353): this CTA's run of the m-block goes to workspace slot mb + blockIdx.x (runs ordered by (m-block, CTA)
                            // form a staircase, so the index is unique per run); lora_partials_reduce_kernel, launched behind this grid, adds an
                            // m-block's runs in CTA order -- lora_act_out needs no zero-fill and every launch gives the same bits.
                            float *mine = p.ws_partial + (static_cast<size_t>(mb + blockIdx.x) * BM + row) * p.Rdp;
                            for (int c16 = h; c16 * 16 < p.Rdp; c16 += H) {
                                uint32_t d2[16];
                                tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + C::kTmemLd + c16 * 16, d2);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 16; i += 4)
                                    __stcg(reinterpret_cast<float4 *>(mine + c16 * 16 + i), make_float4(__uint_as_float(d2[i]), __uint_as_float(d2[i + 1]),
                                                                                                        __uint_as_float(d2[i + 2]), __uint_as_float(d2[i + 3])));
                            }
                            tc_fence_before_sync();
                        } else {
                            // two groups: each adds half of the 16-rank column blocks
                            for (int c16 = h; c16 * 16 < p.Rdp; c16 += H) {
                                uint32_t d2[16];
                                tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + C::kTmemLd + c16 * 16, d2);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 16; i++)
                                    if (c16 * 16 + i < p.R_down && !(p.debug & 16)) atomicAdd(dst + c16 * 16 + i, __uint_as_float(d2[i]));
                            }
                            tc_fence_before_sync();
                        }
                    }
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
        // ============================ INT4 -> hT converter warps ===================================
        if constexpr (!FP4) {
            setmaxnreg_dec<48>();
            const int ct = threadIdx.x - kConvWarp0 * 32;
            PipeState pst, cst;
            long long t_cfull = 0, t_cempty = 0, t_conv = 0;
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
            for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
                for (int kb = 0; kb < num_kblocks; kb++) {
                    // (no cycle counters here: 768 threads x 4 clock reads per k-block showed up in the ncu source view)
                    mbar_wait(&s.full[pst.idx], pst.phase);
                    mbar_wait(&s.cempty[cst.idx], cst.phase ^ 1);
                    if (!(p.debug & 2)) {
                        if (ct < 2 * BM)
                            convert_unit<hT>(smem_u32(s.a[pst.idx]), smem_u32(s.a_cv[cst.idx]), ct, smem_u32(s.sa[pst.idx]), offA, p.debug & 1);
                        else if (ct - 2 * BM < 2 * BN)
                            convert_unit<hT>(smem_u32(s.b[pst.idx]), smem_u32(s.b_cv[cst.idx]), ct - 2 * BM, smem_u32(s.sb[pst.idx]), offB,
                                             p.debug & 1);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {  // one arrival per warp: 512 per-thread arrivals per k-block serialise on one word
                        mbar_arrive(&s.cfull[cst.idx]);
                        mbar_arrive(&s.empty[pst.idx]);
                    }
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

    // ---- teardown -------------------------------------------------------------------------------
    tc_fence_before_sync();
    __syncthreads();
    if (p.prof && threadIdx.x == 0) {
        p.prof[blockIdx.x * 16 + 9] = clock64() - t_kernel0;
        p.prof[blockIdx.x * 16 + 12] = t_setup;
    }
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// lora_act_out[m-block rows][R] = sum over the CTAs whose tile range touched the m-block of the partial projections the fused quantise epilogue left
// in the workspace (slot = m-block + CTA), in a FIXED order: four lanes share one (row, 4 ranks) output, lane p adds the p-th quarter of the runs
// in CTA order and the quarters are combined as (q0 + q1) + (q2 + q3) -- identical bits on every launch, a serial chain of ~19 adds instead of ~74
// at FLUX's text-stream shape (r02h launch list: 28.9 us cold for 2.4 MB of partials with one thread per output).  grid (m-blocks, 128 / kReduceRows).
constexpr int kReduceRows = 16;
constexpr int kReduceThreads = 512;
__global__ void __launch_bounds__(kReduceThreads) lora_partials_reduce_kernel(const float *__restrict__ ws, float *__restrict__ out, int R, int num_n_blocks,
                                                                              int num_tiles, int gemm_grid) {
    ptx::griddep_launch_dependents();
    ptx::griddep_wait();   // the partials are the GEMM's output
    const int mb = blockIdx.x;
    // CTA b of the GEMM owns tiles [tiles * b / grid, tiles * (b + 1) / grid): owner(t) = ceil((t + 1) * grid / tiles) - 1
    const long long t_first = static_cast<long long>(mb) * num_n_blocks, t_last = t_first + num_n_blocks - 1;
    const int b_first = static_cast<int>(((t_first + 1) * gemm_grid + num_tiles - 1) / num_tiles) - 1;
    const int b_last = static_cast<int>(((t_last + 1) * gemm_grid + num_tiles - 1) / num_tiles) - 1;
    const int quarter = (b_last - b_first + 4) / 4;
    const int part = threadIdx.x & 3;
    const int my_first = b_first + part * quarter, my_last = min(b_last, my_first + quarter - 1);
    const int r4 = R / 4;   // (R is a multiple of 16: the 8 (row, 4 ranks) outputs of a warp are all inside or all outside the range below)
    for (int idx = threadIdx.x >> 2; idx < kReduceRows * r4; idx += kReduceThreads / 4) {
        const int row = blockIdx.y * kReduceRows + idx / r4, c4 = idx % r4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int b = my_first; b <= my_last; b++) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(ws + (static_cast<size_t>(mb + b) * 128 + row) * R) + c4);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
#pragma unroll
        for (int sh = 1; sh <= 2; sh <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sh);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sh);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, sh);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, sh);
        }
        if (part == 0) *(reinterpret_cast<float4 *>(out + (static_cast<size_t>(mb) * 128 + row) * R) + c4) = acc;
    }
}

template <bool FP4, typename hT, int BN, int EPI>
int launch(const nb200_gemm_args &a, cudaStream_t stream) {
    using C = Cfg<FP4, BN, EPI>;
    using S = Smem<FP4, BN, EPI>;
    static_assert(sizeof(S) + 1024 <= 232448, "shared memory budget (227 KB per CTA)");
    CUtensorMap tm_act, tm_wgt, tm_out, tm_ld;
    const CUtensorMapSwizzle in_swz =
        FP4 ? (C::kK64 == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B) : CU_TENSOR_MAP_SWIZZLE_NONE;
    const uint32_t in_box = 32 * C::kK64;
    int rc = make_map_2d(&tm_act, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.act, a.K / 2, a.Mp, a.K / 2, in_box, BM, in_swz);
    if (rc) return rc;
    rc = make_map_2d(&tm_wgt, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.wgt, a.K / 2, a.N, a.K / 2, in_box, BN, in_swz);
    if (rc) return rc;
    const CUtensorMapDataType odt =
        HalfTraits<hT>::kIsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (a.out != nullptr) {
        rc = make_map_2d(&tm_out, odt, a.out, a.N_out, a.M_out, static_cast<uint64_t>(a.N_out) * 2, 64, BM,
                         CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    } else {
        tm_out = tm_act;  // never dereferenced
    }
    const int rdp = a.R_down;
    if (EPI == EPI_QUANT && a.R_down > 0) {
        rc = make_map_2d(&tm_ld, odt, a.lora_down_next, a.N, rdp, static_cast<uint64_t>(a.N) * 2, 64, rdp,
                         CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        // without a workspace the partial projections are ADDED to lora_act_out with fp32 atomics: zero it first (as the reference does, launch_impl:252)
        if (a.workspace == nullptr || a.workspace_bytes < nb200_gemm_workspace_bytes(a.Mp, a.R_down))
            NB200_CUDA_CHECK(cudaMemsetAsync(a.lora_act_out, 0, static_cast<size_t>(a.Mp) * a.R_down * sizeof(float), stream));
    } else {
        tm_ld = tm_act;
    }

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
    p.prof = static_cast<long long *>(a.prof);
    p.tile_contig = (EPI == EPI_QUANT && a.R_down > 0) ? 1 : 0;
    static const int dbg = getenv("NB200_GEMM_DEBUG") ? atoi(getenv("NB200_GEMM_DEBUG")) : 0;
    p.debug = dbg;
    p.out = a.out;
    p.qout = static_cast<uint8_t *>(a.qout);
    p.oscales_out = a.oscales;
    p.smooth_next = a.smooth_next;
    p.lora_act_out = a.lora_act_out;
    p.R_down = a.R_down;
    p.Rdp = rdp;
    p.ws_partial = nullptr;
    if (EPI == EPI_QUANT && a.R_down > 0 && a.workspace != nullptr && a.workspace_bytes >= nb200_gemm_workspace_bytes(a.Mp, a.R_down))
        p.ws_partial = static_cast<float *>(a.workspace);
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
    p.out_vk = a.out_vk;
    p.vk_tokens = a.vk_tokens;
    p.vk_heads = a.N / 96;
    if (EPI == EPI_LITELA)   // the reference zero-fills out_vk inside the op as well (launch_impl:336)
        NB200_CUDA_CHECK(cudaMemsetAsync(a.out_vk, 0, static_cast<size_t>(a.Mp / a.vk_tokens) * p.vk_heads * 33 * 32 * sizeof(float), stream));
    for (int i = 0; i < NB200_MAX_LORA_SCALES; i++) p.lora_scales[i] = a.lora_scales[i];

    int num_sms_dev = 0;
    if (int rc2 = current_device_sms(&num_sms_dev)) return rc2;
    const int num_sms = a.num_sms > 0 ? a.num_sms : num_sms_dev;
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    if (grid > kWsMaxCtas) p.ws_partial = nullptr;   // (the workspace is sized for kWsMaxCtas runs; p.ws_partial == nullptr selects the atomics)
    if (EPI == EPI_QUANT && a.R_down > 0 && p.ws_partial == nullptr && a.workspace != nullptr && a.workspace_bytes >= nb200_gemm_workspace_bytes(a.Mp, a.R_down))
        NB200_CUDA_CHECK(cudaMemsetAsync(a.lora_act_out, 0, static_cast<size_t>(a.Mp) * a.R_down * sizeof(float), stream));
    const size_t smem_bytes = sizeof(S) + 1024;
    auto kern = gemm_w4a4_kernel<FP4, hT, BN, EPI>;
    if (int rc2 = set_max_smem_once(reinterpret_cast<const void *>(kern), smem_bytes)) return rc2;
    LaunchCfg lc(dim3(grid), dim3(C::kThreads), smem_bytes, stream);
    NB200_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, tm_act, tm_wgt, tm_out, tm_ld, p));
    count_launch();
    NB200_CUDA_CHECK(cudaGetLastError());
    if (p.ws_partial != nullptr) {   // the fixed-order sum of the per-CTA partial projections
        LaunchCfg lr(dim3(a.Mp / BM, BM / kReduceRows), dim3(kReduceThreads), 0, stream);
        NB200_CUDA_CHECK(cudaLaunchKernelEx(&lr.cfg, lora_partials_reduce_kernel, static_cast<const float *>(p.ws_partial), a.lora_act_out, a.R_down, p.num_n_blocks,
                                            p.num_tiles, grid));
        count_launch();
    }
    return NB200_OK;
}

template <bool FP4, typename hT>
int launch_bn(const nb200_gemm_args &a, cudaStream_t stream) {
    if (a.out_vk != nullptr) return launch<FP4, hT, 128, EPI_LITELA>(a, stream);   // SANA linear attention: two heads of [K 32 | V 32] per 128-wide tile
    if (a.qout != nullptr) {
        // INT4: 256-wide tiles halve the converter work per MMA cycle (the INT4 limiter); needs the next layer's rank to fit
        // the smaller smem / TMEM budget.  block_n = 128 forces the narrow tile.  (NVFP4: the 256-wide variant does not fit smem.)
        if constexpr (!FP4) {
            if (a.block_n != 128 && a.N % 256 == 0 && a.R_down <= 32 && (a.block_n == 256 || (a.Mp / BM) * (a.N / 256) >= 96))
                return launch<FP4, hT, 256, EPI_QUANT>(a, stream);
        }
        return launch<FP4, hT, 128, EPI_QUANT>(a, stream);
    }
    // NVFP4 on clusters (gemm_nvfp4_cluster.cu): 256x256 pair tiles, two pairs per cluster sharing A by TMA multicast.
    // block_n = 1024 / 2048 force one / two pairs per cluster; auto picks it whenever there are enough pair tiles to
    // fill the chip (the 256-token text stream keeps the narrow single-CTA tiles: more CTAs busy).
    if constexpr (FP4) {
        if (a.block_n == 1024 || a.block_n == 2048) return gemm_nvfp4_cluster_dispatch(a, stream, a.block_n / 1024);
        // measured on one box (tools/gemm_ablate.py): one pair per cluster == the r01 pair kernel on the plain
        // epilogue and it carries the RoPE epilogue on 256-wide tiles; two pairs per cluster (A multicast) is NOT faster -- the bound
        // is what each SM has to receive, which multicast does not change -- and only 33 clusters of 4 fit on the 148 SMs
        if (a.block_n == 0 && a.N % 256 == 0 && (a.Mp / 256) * (a.N / 256) >= 64) return gemm_nvfp4_cluster_dispatch(a, stream, 1);
    }
    if (a.rotary_emb != nullptr) return launch<FP4, hT, 128, EPI_ROPE>(a, stream);
    if (a.block_n == 512) {  // CTA pairs (cta_group::2), 256 x 256 tiles
        if (a.N % 256 != 0) return fail(NB200_ERR_INVALID_ARGUMENT, "the CTA-pair kernel needs N % 256 == 0");
        return gemm_w4a4_2cta_dispatch(a, stream);
    }
    int bn = a.block_n;
    // measured (tools/op_sweep.py, profiles/): the CTA-pair kernel wins once the main loop is long enough to
    // amortise its fully exposed epilogue hand-off (M=4352, K=12288: 3.83 vs 3.61 PFLOP/s); shorter K loses
    if (bn == 0 && FP4 && a.K >= 8192 && a.N % 256 == 0 && a.Mp >= 2048) return gemm_w4a4_2cta_dispatch(a, stream);
    if (bn == 0) bn = (a.N % 256 == 0 && (a.Mp / BM) * (a.N / 256) >= 96) ? 256 : 128;
    if (bn == 256 && a.N % 256 == 0) return launch<FP4, hT, 256, EPI_DEFAULT>(a, stream);
    if (bn == 128) return launch<FP4, hT, 128, EPI_DEFAULT>(a, stream);
    return fail(NB200_ERR_INVALID_ARGUMENT, "block_n must be 0, 128, 256 (single CTA) or 512 (CTA pair) and divide N");
}

}  // namespace
}  // namespace nb200

// one [128][R_down] fp32 slot per (m-block, CTA) run
extern "C" __attribute__((visibility("default"))) long long nb200_gemm_workspace_bytes(int Mp, int R_down) {
    if (Mp <= 0 || R_down <= 0) return 0;
    return (static_cast<long long>(Mp / 128) + nb200::kWsMaxCtas) * 128 * R_down * 4;
}

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
    NB200_REQUIRE(a->out != nullptr || a->qout != nullptr || a->out_q != nullptr, "out, qout or out_q must be non-NULL");
    if (a->qout != nullptr) {
        // fc1 -> GELU -> (lora_down of fc2) -> quantise for fc2   (launch_impl:282-310)
        NB200_REQUIRE(a->oscales && a->smooth_next, "qout needs oscales and smooth_next");
        NB200_REQUIRE(a->rotary_emb == nullptr, "qout and rotary_emb are exclusive");
        NB200_REQUIRE(a->R_down >= 0 && a->R_down % 16 == 0, "R_down must be a multiple of 16");
        if (a->R_down > kMaxRdp)
            return fail(NB200_ERR_UNSUPPORTED, "fused next-layer low-rank down projection supports rank <= 128");
        NB200_REQUIRE(a->R_down == 0 || (a->lora_down_next && a->lora_act_out), "lora_down_next and lora_act_out go together");
        NB200_REQUIRE((reinterpret_cast<uintptr_t>(a->qout) & 15) == 0, "qout must be 16-byte aligned");
    }
    if (a->rotary_emb != nullptr) {
        NB200_REQUIRE(a->norm_q && a->norm_k && (a->out || a->out_q), "rotary_emb needs norm_q, norm_k and out (or out_q/k/v)");
        NB200_REQUIRE(a->N % 384 == 0, "RMSNorm+RoPE epilogue: N must be 3 * heads * 128");
    }
    if (a->out_q != nullptr || a->out_k != nullptr || a->out_v != nullptr) {
        NB200_REQUIRE(a->rotary_emb != nullptr, "out_q/out_k/out_v are outputs of the RMSNorm+RoPE epilogue (launch_impl:376)");
        NB200_REQUIRE(a->out_q && a->out_k && a->out_v, "out_q, out_k and out_v go together");
        // (the reference writes either `out` or out_q/k/v, launch_impl:376-405; here `out` next to out_q/k/v is an optional [Mp, N] SCRATCH the
        //  launcher may overwrite: it lets the NVFP4 cluster route run its plain epilogue + the RMSNorm / RoPE / pack kernel of rope.cu)
        NB200_REQUIRE(a->out == nullptr || (a->M_out == a->Mp && a->N_out == a->N), "with out_q/k/v, out is a scratch of exactly [Mp, N]");
        NB200_REQUIRE(a->attn_tokens >= 0 && a->attn_tokens <= a->Mp, "attn_tokens must be in [0, Mp]");
        NB200_REQUIRE(a->stride_head_q >= static_cast<long long>(a->Mp) * 128 && a->stride_head_k >= static_cast<long long>(a->Mp) * 128 &&
                          a->stride_head_v >= static_cast<long long>(a->Mp) * 128,
                      "head pitch must cover Mp rows of 128");
        NB200_REQUIRE(((reinterpret_cast<uintptr_t>(a->out_q) | reinterpret_cast<uintptr_t>(a->out_k) | reinterpret_cast<uintptr_t>(a->out_v)) & 15) == 0 &&
                          a->stride_head_q % 8 == 0 && a->stride_head_k % 8 == 0 && a->stride_head_v % 8 == 0,
                      "out_q/k/v must be 16-byte aligned");
    }
    if (a->out_vk != nullptr) {
        // EpilogueLiteLA (launch_impl:311-346): out = relu(Q) [Mp, N/3], out_vk [Mp / vk_tokens][N / 96][33][32] f32
        NB200_REQUIRE(a->out != nullptr && a->qout == nullptr && a->rotary_emb == nullptr && a->mid_act == NB200_ACT_NONE,
                      "out_vk goes with out (= out_linearattn) only");
        NB200_REQUIRE(a->N % 384 == 0, "LiteLA epilogue: N / 3 must be a multiple of 128 (numBlocksN % 3 == 0, epilogues.cuh:585)");
        NB200_REQUIRE(a->vk_tokens > 0 && a->vk_tokens % 128 == 0 && a->Mp % a->vk_tokens == 0, "vk_tokens must be a multiple of 128 that divides Mp");
        NB200_REQUIRE(a->M_out == a->Mp && a->N_out == a->N / 3, "LiteLA epilogue: out is [Mp, N / 3]");
    }
    if (a->out != nullptr) {
        NB200_REQUIRE(a->M_out > 0 && a->M_out <= a->Mp && a->Mp - a->M_out < 256, "M_out must be in (Mp-256, Mp]");
        NB200_REQUIRE(a->out_vk != nullptr || (a->N_out > 0 && a->N_out <= a->N && a->N - a->N_out < 128), "N_out must be in (N-128, N]");
        NB200_REQUIRE(a->N_out % 8 == 0, "out row pitch must be a multiple of 16 bytes (TMA store)");
        NB200_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "out must be 16-byte aligned");
    }
    NB200_REQUIRE((reinterpret_cast<uintptr_t>(a->act) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->wgt) & 15) == 0,
                  "act/wgt must be 16-byte aligned");
    NB200_REQUIRE(a->fp4 || !a->cscale, "INT4 path has no per-channel scale (alpha == 1, launch_impl:107)");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (a->fp4) {
        NB200_REQUIRE(!a->act_unsigned, "act_unsigned is INT4 only");
        return a->dtype == NB200_BF16 ? launch_bn<true, __nv_bfloat16>(*a, stream) : launch_bn<true, __half>(*a, stream);
    }
    return a->dtype == NB200_BF16 ? launch_bn<false, __nv_bfloat16>(*a, stream) : launch_bn<false, __half>(*a, stream);
}
