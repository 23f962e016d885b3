// Do tcgen05.mma (operands resident in shared memory) and TMA loads into OTHER shared memory of the same SM slow each other down?
// (r02: the NVFP4 GEMM runs at ~930 clk per k-stage; its MMAs alone need 512 clk, its loads alone ~670 clk -- tools/gemm_ablate.py.)
//
// One CTA (or CTA pair) per SM.  Warp 0: back-to-back NVFP4 MMAs on 2 resident operand stages (as tools/ubench/mma_peak.cu).
// Warp 1: a ring of 3 bulk copies of `load_kb` KB each from a 512 MB global buffer (L2 misses and hits mixed like a GEMM's
// operand stream is not the point: the buffer slice per SM is re-read, so it is L2 resident after the first pass).
// Modes: 1 = MMAs only, 2 = loads only, 3 = both; 6 / 7 = the same loads as 2-D tensor-map boxes instead of 1-D bulk copies: per
// 38 KB "stage" two boxes of 128 rows x 128 bytes with SWIZZLE_128B out of a row-major [rows, 1536 B] matrix (what the GEMM's A / B
// operand loads look like) plus 6 KB of 1-D bulk copy (its scale factors).  Reports clk per MMA and load bytes per clk per SM.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I nunchaku_b200/csrc tools/ubench/mma_tma_interference.cu -o tools/ubench/_bin/mma_tma_interference
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>

#include "ptx.cuh"

using namespace nb200::ptx;

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kMmaStages = 2;
constexpr int kLoadSlots = 3;
constexpr int kLoadBytes = 38 * 1024;

struct alignas(1024) Smem {
    alignas(1024) uint8_t a[kMmaStages][128 * 128];
    alignas(1024) uint8_t b[kMmaStages][256 * 128];
    alignas(1024) uint8_t ld[kLoadSlots][kLoadBytes];
    alignas(128) uint8_t sf[16 * 512];
    uint64_t done, ldbar[kLoadSlots];
    uint32_t tmem_base;
};

template <bool CG2>
__global__ void __launch_bounds__(128, 1) kern(int iters, int mode, const uint8_t *src, size_t per_sm_bytes, long long *out,
                                               const __grid_constant__ CUtensorMap tm, int rows_per_sm) {
    extern __shared__ uint8_t raw[];
    Smem &s = *reinterpret_cast<Smem *>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5;
    const bool leader = !CG2 || cluster_ctarank() == 0;
    for (int i = threadIdx.x; i < (int)(sizeof(s.a) + sizeof(s.b)); i += blockDim.x) {
        uint32_t h = (i + blockIdx.x * 7919u) * 2654435761u;
        reinterpret_cast<uint8_t *>(s.a)[i] = static_cast<uint8_t>((h ^ (h >> 15)) >> 8);
    }
    for (int i = threadIdx.x; i < (int)sizeof(s.sf); i += blockDim.x) s.sf[i] = 0x30 + (i & 7);
    if (threadIdx.x == 0) {
        mbar_init(&s.done, 1);
        for (int i = 0; i < kLoadSlots; i++) mbar_init(&s.ldbar[i], 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    if (warp == 2) {
        if (CG2) tmem_alloc_cg2<512>(&s.tmem_base); else tmem_alloc<512>(&s.tmem_base);
    }
    tc_fence_before_sync();
    if (CG2) cluster_sync(); else __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    constexpr uint32_t M = CG2 ? 256 : 128, N = 256;
    if (warp == 0 && leader && (mode & 1) && elect_one()) {
        const uint32_t sfa = tmem + 256, sfb = tmem + 272;
        for (int j = 0; j < 4; j++) {
            const uint64_t d0 = make_smem_desc(smem_u32(s.sf + j * 512), 0, 128, kLayoutNoSwizzle);
            const uint64_t d1 = make_smem_desc(smem_u32(s.sf + (4 + j) * 512), 0, 128, kLayoutNoSwizzle);
            const uint64_t d2 = make_smem_desc(smem_u32(s.sf + (8 + j) * 512), 0, 128, kLayoutNoSwizzle);
            if (CG2) { tc_cp_32x128b_warpx4_cg2(sfa + 4 * j, d0); tc_cp_32x128b_warpx4_cg2(sfb + 8 * j, d1); tc_cp_32x128b_warpx4_cg2(sfb + 8 * j + 4, d2); }
            else { tc_cp_32x128b_warpx4(sfa + 4 * j, d0); tc_cp_32x128b_warpx4(sfb + 8 * j, d1); tc_cp_32x128b_warpx4(sfb + 8 * j + 4, d2); }
        }
        constexpr uint32_t idesc = make_idesc_nvf4(M, N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            const int st = it & (kMmaStages - 1);
            const uint32_t a_addr = smem_u32(s.a[st]), b_addr = smem_u32(s.b[st]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t ad = make_sw128_kmajor_desc(a_addr + j * 32), bd = make_sw128_kmajor_desc(b_addr + j * 32);
                if (CG2) tc_mma_nvf4_cg2(tmem, ad, bd, idesc, sfa + 4 * j, sfb + 8 * j, (it | j) != 0);
                else tc_mma_nvf4(tmem, ad, bd, idesc, sfa + 4 * j, sfb + 8 * j, (it | j) != 0);
            }
        }
        if (CG2) tc_commit_cg2(&s.done, 1); else tc_commit(&s.done);
        mbar_wait(&s.done, 0);
        out[blockIdx.x * 4 + 0] = clock64() - t0;
    }
    if (warp == 1 && (mode & 2) && elect_one()) {
        if (mode & 4) prefetch_tensormap(&tm);
        // as many bytes as the MMA loop's duration would need at one 38 KB stage per 4 MMAs: iters stages
        const uint8_t *base = src + static_cast<size_t>(blockIdx.x) * per_sm_bytes;
        const long long t0 = clock64();
        uint32_t phase[kLoadSlots] = {0, 0, 0};
        size_t off = 0;
        for (int it = 0; it < iters; it++) {
            const int sl = it % kLoadSlots;
            if (it >= kLoadSlots) {
                mbar_wait(&s.ldbar[sl], phase[sl]);
                phase[sl] ^= 1;
            }
            mbar_expect_tx(&s.ldbar[sl], kLoadBytes);
            if (mode & 4) {
                // matrix [148 * rows_per_sm rows][1536 B]: this SM's rows, walking along k (12 boxes of 128 B) then down the rows
                const int kb = it % 12, rb = (it / 12) % (rows_per_sm / 256);
                const int row0 = blockIdx.x * rows_per_sm + rb * 256;
                tma_load_2d(s.ld[sl], &tm, &s.ldbar[sl], kb * 128, row0);
                tma_load_2d(s.ld[sl] + 16384, &tm, &s.ldbar[sl], kb * 128, row0 + 128);
                bulk_load(s.ld[sl] + 32768, base + (off % (per_sm_bytes - 6144)) / 16 * 16, 6144, &s.ldbar[sl]);
            } else {
                bulk_load(s.ld[sl], base + off, kLoadBytes, &s.ldbar[sl]);
            }
            off += kLoadBytes;
            if (off + kLoadBytes > per_sm_bytes) off = 0;
        }
        for (int k = 0; k < kLoadSlots && k < iters; k++) {
            const int sl = (iters - 1 - k) % kLoadSlots;
            mbar_wait(&s.ldbar[sl], phase[sl]);
        }
        out[blockIdx.x * 4 + 1] = clock64() - t0;
    }
    tc_fence_before_sync();
    if (CG2) cluster_sync(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        if (CG2) tmem_dealloc_cg2<512>(tmem); else tmem_dealloc<512>(tmem);
    }
}

template <bool CG2>
void run(int num_sms, int iters, const uint8_t *src, size_t per_sm, const CUtensorMap &tm, int rows_per_sm) {
    auto k = kern<CG2>;
    const size_t smem = sizeof(Smem) + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long *out;
    cudaMalloc(&out, num_sms * 4 * sizeof(long long));
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
    for (int mode : {1, 2, 3, 6, 7}) {
        cudaMemset(out, 0, num_sms * 4 * sizeof(long long));
        for (int rep = 0; rep < 3; rep++) cudaLaunchKernelEx(&cfg, k, iters, mode, src, per_sm, out, tm, rows_per_sm);
        cudaDeviceSynchronize();
        if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); exit(1); }
        std::vector<long long> h(num_sms * 4);
        cudaMemcpy(h.data(), out, num_sms * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
        double cm = 0, cl = 0; int nm = 0, nl = 0;
        for (int i = 0; i < grid; i++) {
            if (h[i * 4]) { cm += h[i * 4]; nm++; }
            if (h[i * 4 + 1]) { cl += h[i * 4 + 1]; nl++; }
        }
        printf(", \"%s_mode%d\": {\"clk_per_mma\": %.1f, \"load_bytes_per_clk_per_sm\": %.1f}", CG2 ? "cg2" : "cg1", mode,
               nm ? cm / nm / (iters * 4.0) : 0.0, nl ? (double)iters * kLoadBytes / (cl / nl) : 0.0);
        fflush(stdout);
    }
    cudaFree(out);
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    int dev = 0, num_sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    // per SM 512 rows x 1536 B = 768 KB (L2 resident after the first pass): 148 x 768 KB = 114 MB
    const int rows_per_sm = 512;
    const size_t per_sm_bytes = static_cast<size_t>(rows_per_sm) * 1536;
    uint8_t *src;
    cudaMalloc(&src, per_sm_bytes * num_sms);
    cudaMemset(src, 0x5a, per_sm_bytes * num_sms);
    CUtensorMap tm;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        const cuuint64_t dims[2] = {1536, static_cast<cuuint64_t>(rows_per_sm) * num_sms};
        const cuuint64_t strides[1] = {1536};
        const cuuint32_t box[2] = {128, 128};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = reinterpret_cast<EncodeTiledFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed %d\n", (int)r); return 1; }
    }
    printf("{\"what\": \"tools/ubench/mma_tma_interference.cu: NVFP4 MMAs on resident operands (mode 1), 38 KB loads per stage from an L2-resident buffer into a 3-slot ring as ONE 1-D bulk copy (mode 2) or as two 128x128B SWIZZLE_128B tensor-map boxes + 6 KB bulk (mode 6), MMAs + loads together (modes 3, 7)\", \"iters\": %d", iters);
    run<false>(num_sms, iters, src, per_sm_bytes, tm, rows_per_sm);
    run<true>(num_sms, iters, src, per_sm_bytes, tm, rows_per_sm);
    printf("}\n");
    return 0;
}
