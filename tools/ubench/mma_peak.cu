// Measured tcgen05 tensor-pipe peaks on the GPU it runs on (roofline denominators for bench.py; VERDICT r01 item 5:
// "FP4/INT8/BF16 tcgen05 peak from a committed microbench instead of '4x'").
//
// One CTA per SM (or one CTA pair per TPC for cta_group::2).  Operand tiles sit in shared memory (pseudo-random bytes,
// no global loads in the timed region), one thread issues back-to-back tcgen05.mma into a TMEM accumulator and commits
// once; the kernel is timed with CUDA events, the MMA loop with clock64 / globaltimer inside.
//   kind::f16 (bf16)      M128|256 x N256 x K16
//   kind::i8  (s8 x s8)   M128|256 x N256 x K32
//   kind::f8f6f4 (e4m3)   M128|256 x N256 x K32
//   kind::mxf4nvf4.block_scale.scale_vec::4X (e2m1, ue4m3)  M128|256 x N256 x K64
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I nunchaku_b200/csrc tools/ubench/mma_peak.cu -o tools/ubench/_bin/mma_peak
//   tools/ubench/_bin/mma_peak [iters] [seconds_sustained]   -> one JSON object on stdout
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace nb200::ptx;

enum Kind { KIND_BF16 = 0, KIND_I8 = 1, KIND_FP8 = 2, KIND_NVF4 = 3 };

__host__ __device__ constexpr uint32_t idesc_dense(uint32_t cfmt, uint32_t afmt, uint32_t bfmt, uint32_t M, uint32_t N) {
    return (cfmt << 4) | (afmt << 7) | (bfmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, bool cg2) {
    if (cg2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
                     "r"(idesc), "r"(acc)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
                     "r"(idesc), "r"(acc)
                     : "memory");
}
__device__ __forceinline__ void mma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, bool cg2) {
    if (cg2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a),
                     "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a),
                     "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
}

constexpr int kStages = 4;
constexpr int kABytes = 128 * 128, kBBytes = 256 * 128;

struct alignas(1024) Smem {
    alignas(1024) uint8_t a[kStages][kABytes];
    alignas(1024) uint8_t b[kStages][kBBytes];   // cg2: each CTA holds 128 of the 256 rows (first half used)
    alignas(128) uint8_t sf[16 * 512];
    uint64_t done;
    uint32_t tmem_base;
};

// cp_mode (NVF4 only): 0 = scale factors resident in TMEM; 1 = 12 tcgen05.cp.32x128b.warpx4 per 4 MMAs into alternating TMEM sets
// (what the GEMM main loop does); 2 = the same into ONE set (write-after-read hazard against the running MMAs); 3 = the copies alone
template <int KIND, bool CG2>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters, long long *cycles, long long *nanos, int cp_mode) {
    extern __shared__ uint8_t raw[];
    Smem &s = *reinterpret_cast<Smem *>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5;
    const bool leader = !CG2 || cluster_ctarank() == 0;
    // pseudo-random operand bytes (data toggling matters for power / clocks); fp8 / nvf4 codes avoid NaN patterns
    for (int i = threadIdx.x; i < (int)(sizeof(s.a) + sizeof(s.b)); i += blockDim.x) {
        uint32_t h = (i + blockIdx.x * 7919u) * 2654435761u;
        h ^= h >> 15;
        uint8_t v = static_cast<uint8_t>(h >> 8);
        if (KIND == KIND_FP8 && (v & 0x7F) == 0x7F) v &= 0xF7;
        reinterpret_cast<uint8_t *>(s.a)[i] = v;
    }
    for (int i = threadIdx.x; i < (int)sizeof(s.sf); i += blockDim.x) s.sf[i] = 0x30 + (i & 7);   // ue4m3 0.5 .. 0.94
    if (threadIdx.x == 0) {
        mbar_init(&s.done, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    if (warp == 2) {
        if (CG2)
            tmem_alloc_cg2<512>(&s.tmem_base);
        else
            tmem_alloc<512>(&s.tmem_base);
    }
    tc_fence_before_sync();
    if (CG2)
        cluster_sync();
    else
        __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    constexpr uint32_t M = CG2 ? 256 : 128, N = 256;
    if (warp == 0 && leader && elect_one()) {
        const uint32_t sfa = tmem + 256, sfb = tmem + 272;
        if (KIND == KIND_NVF4) {
            for (int j = 0; j < 4; j++) {
                const uint64_t d0 = make_smem_desc(smem_u32(s.sf + j * 512), 0, 128, kLayoutNoSwizzle);
                const uint64_t d1 = make_smem_desc(smem_u32(s.sf + (4 + j) * 512), 0, 128, kLayoutNoSwizzle);
                const uint64_t d2 = make_smem_desc(smem_u32(s.sf + (8 + j) * 512), 0, 128, kLayoutNoSwizzle);
                if (CG2) {
                    tc_cp_32x128b_warpx4_cg2(sfa + 4 * j, d0);
                    tc_cp_32x128b_warpx4_cg2(sfb + 8 * j, d1);
                    tc_cp_32x128b_warpx4_cg2(sfb + 8 * j + 4, d2);
                } else {
                    tc_cp_32x128b_warpx4(sfa + 4 * j, d0);
                    tc_cp_32x128b_warpx4(sfb + 8 * j, d1);
                    tc_cp_32x128b_warpx4(sfb + 8 * j + 4, d2);
                }
            }
        }
        constexpr uint32_t idesc = KIND == KIND_BF16 ? idesc_dense(1, 1, 1, M, N)
                                   : KIND == KIND_I8 ? idesc_dense(2, 1, 1, M, N)
                                   : KIND == KIND_FP8 ? idesc_dense(1, 0, 0, M, N)
                                                      : make_idesc_nvf4(M, N);
        long long t0 = clock64();
        unsigned long long n0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
        for (int it = 0; it < iters; it++) {
            const int st = it & (kStages - 1);
            const uint32_t a_addr = smem_u32(s.a[st]), b_addr = smem_u32(s.b[st]);
            uint32_t sfa_it = sfa, sfb_it = sfb;
            if (KIND == KIND_NVF4 && cp_mode != 0) {
                const uint32_t set = (cp_mode == 1 ? (it & 1) : 0) * 48;
                sfa_it = sfa + set;
                sfb_it = sfb + set;   // sets: [256, 304) and [304, 352): sfa 16 columns, sfb 32 columns each
                for (int j = 0; j < 4; j++) {
                    const uint64_t d0 = make_smem_desc(smem_u32(s.sf + j * 512), 0, 128, kLayoutNoSwizzle);
                    const uint64_t d1 = make_smem_desc(smem_u32(s.sf + (4 + j) * 512), 0, 128, kLayoutNoSwizzle);
                    const uint64_t d2 = make_smem_desc(smem_u32(s.sf + (8 + j) * 512), 0, 128, kLayoutNoSwizzle);
                    if (CG2) {
                        tc_cp_32x128b_warpx4_cg2(sfa_it + 4 * j, d0);
                        tc_cp_32x128b_warpx4_cg2(sfb_it + 8 * j, d1);
                        tc_cp_32x128b_warpx4_cg2(sfb_it + 8 * j + 4, d2);
                    } else {
                        tc_cp_32x128b_warpx4(sfa_it + 4 * j, d0);
                        tc_cp_32x128b_warpx4(sfb_it + 8 * j, d1);
                        tc_cp_32x128b_warpx4(sfb_it + 8 * j + 4, d2);
                    }
                }
                if (cp_mode == 3) continue;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {   // 4 x 32 bytes of K per 128-byte swizzled row
                const uint64_t ad = make_sw128_kmajor_desc(a_addr + j * 32), bd = make_sw128_kmajor_desc(b_addr + j * 32);
                const uint32_t acc = (it | j) != 0;
                if (KIND == KIND_BF16) {
                    if (CG2) tc_mma_f16_cg2(tmem, ad, bd, idesc, acc); else tc_mma_f16(tmem, ad, bd, idesc, acc);
                } else if (KIND == KIND_I8) {
                    mma_i8(tmem, ad, bd, idesc, acc, CG2);
                } else if (KIND == KIND_FP8) {
                    mma_f8(tmem, ad, bd, idesc, acc, CG2);
                } else {
                    if (CG2) tc_mma_nvf4_cg2(tmem, ad, bd, idesc, sfa_it + 4 * j, sfb_it + 8 * j, acc);
                    else tc_mma_nvf4(tmem, ad, bd, idesc, sfa_it + 4 * j, sfb_it + 8 * j, acc);
                }
            }
        }
        if (CG2) tc_commit_cg2(&s.done, 1); else tc_commit(&s.done);
        mbar_wait(&s.done, 0);
        unsigned long long n1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
        cycles[blockIdx.x] = clock64() - t0;
        nanos[blockIdx.x] = (long long)(n1 - n0);
    }
    tc_fence_before_sync();
    if (CG2)
        cluster_sync();
    else
        __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        if (CG2)
            tmem_dealloc_cg2<512>(tmem);
        else
            tmem_dealloc<512>(tmem);
    }
}

struct Result {
    double tflops_burst, tflops_sustained, clk_per_mma, eff_mhz;
};

template <int KIND, bool CG2>
Result run(int num_sms, int iters, double sustain_s, int cp_mode = 0) {
    auto kern = mma_peak_kernel<KIND, CG2>;
    const size_t smem = sizeof(Smem) + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long *cyc, *ns;
    cudaMalloc(&cyc, num_sms * sizeof(long long));
    cudaMalloc(&ns, num_sms * sizeof(long long));
    cudaLaunchConfig_t cfg = {};
    const int grid = CG2 ? (num_sms / 2) * 2 : num_sms;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG2 ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const double K = KIND == KIND_BF16 ? 16 : KIND == KIND_NVF4 ? 64 : 32;
    const double flop_per_launch = (double)(CG2 ? grid / 2 : grid) * iters * 4.0 * 2.0 * (CG2 ? 256 : 128) * 256 * K;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto once = [&]() {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, kern, iters, cyc, ns, cp_mode);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        return (double)ms;
    };
    for (int i = 0; i < 3; i++) once();
    double best = 1e30;
    for (int i = 0; i < 10; i++) best = std::min(best, once());
    if (cudaGetLastError() != cudaSuccess) {
        fprintf(stderr, "kernel failed\n");
        exit(1);
    }
    std::vector<long long> hc(grid), hn(grid);
    cudaMemcpy(hc.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(hn.data(), ns, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double c = 0, n = 0;
    int cnt = 0;
    for (int i = 0; i < grid; i += (CG2 ? 2 : 1)) {
        c += hc[i];
        n += hn[i];
        cnt++;
    }
    c /= cnt;
    n /= cnt;
    // sustained: back-to-back launches for sustain_s seconds, average rate of the second half
    double t_total = 0, t_half = 0;
    int launches_half = 0;
    while (t_total < sustain_s * 1e3) {
        double ms = once();
        t_total += ms;
        if (t_total > sustain_s * 500.0) {
            t_half += ms;
            launches_half++;
        }
    }
    Result r;
    r.tflops_burst = flop_per_launch / (best * 1e-3) / 1e12;
    r.tflops_sustained = launches_half ? flop_per_launch * launches_half / (t_half * 1e-3) / 1e12 : 0;
    r.clk_per_mma = c / (iters * 4.0);
    r.eff_mhz = c / n * 1e3;
    cudaFree(cyc);
    cudaFree(ns);
    return r;
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4000;
    const double sustain = argc > 2 ? atof(argv[2]) : 2.0;
    int dev = 0, num_sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"iters\": %d, \"how\": \"tools/ubench/mma_peak.cu: smem-resident operands, 1 CTA (or CTA pair) per SM, back-to-back tcgen05.mma; "
           "burst = best of 10 launches, sustained = second half of %.1f s of back-to-back launches\"",
           prop.name, num_sms, iters, sustain);
    const char *names[4] = {"bf16", "i8", "fp8", "nvf4"};
    Result r;
#define RUN(K, C)                                                                                                                                    \
    r = run<K, C>(num_sms, iters, sustain);                                                                                                          \
    printf(", \"%s_%s\": {\"tflops_burst\": %.1f, \"tflops_sustained\": %.1f, \"clk_per_mma\": %.1f, \"sm_mhz_in_loop\": %.0f}", names[K],          \
           C ? "cg2" : "cg1", r.tflops_burst, r.tflops_sustained, r.clk_per_mma, r.eff_mhz);                                                         \
    fflush(stdout);
    RUN(KIND_BF16, false)
    RUN(KIND_BF16, true)
    RUN(KIND_I8, false)
    RUN(KIND_I8, true)
    RUN(KIND_FP8, false)
    RUN(KIND_FP8, true)
    RUN(KIND_NVF4, false)
    RUN(KIND_NVF4, true)
    // the GEMM main loop's scale-factor traffic: 12 tcgen05.cp per 4 MMAs (tflops still count the MMAs only; mode 3 = copies alone,
    // its "clk_per_mma" is the cost of 3 copies)
    const char *cpn[4] = {"", "cp_two_sets", "cp_one_set", "cp_alone"};
    for (int mode = 1; mode <= 3; mode++)
        for (int c = 0; c < 2; c++) {
            r = c ? run<KIND_NVF4, true>(num_sms, iters, 0.3, mode) : run<KIND_NVF4, false>(num_sms, iters, 0.3, mode);
            printf(", \"nvf4_%s_%s\": {\"tflops_burst\": %.1f, \"tflops_sustained\": %.1f, \"clk_per_mma\": %.1f, \"sm_mhz_in_loop\": %.0f}", c ? "cg2" : "cg1", cpn[mode],
                   r.tflops_burst, r.tflops_sustained, r.clk_per_mma, r.eff_mhz);
            fflush(stdout);
        }
    printf("}\n");
    return 0;
}
