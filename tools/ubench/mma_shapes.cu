// Cost of ONE tcgen05.mma kind::f16 (fp16 operands, fp32 accumulate, cta_group::1, M = 128, K = 16) by operand source and N:
// the question behind csrc/attention.cu's tile shapes (r02): what do the attention GEMMs' instructions really cost?
//   SS  K-major A, K-major B            N = 64 / 128 / 256     (S = Q K^T)
//   SS  K-major A, MN-major B           N = 128                (O += P V with P in shared memory: attention v1)
//   TS  A in TMEM, MN-major B           N = 128                (O += P V with P in TMEM: attention v2+)
//   TS  A in TMEM, K-major B            N = 128
// One CTA per SM, operands resident in shared memory / TMEM, one thread issues `iters` back-to-back instructions and commits once.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I nunchaku_b200/csrc tools/ubench/mma_shapes.cu -o tools/ubench/_bin/mma_shapes
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace nb200::ptx;

struct alignas(1024) Smem {
    alignas(1024) uint8_t a[4][128 * 128];
    alignas(1024) uint8_t b[4][256 * 128];
    uint64_t done;
    uint32_t tmem_base;
};

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}

// mode: 0 SS K/K, 1 SS K/MN, 2 TS MN, 3 TS K
__global__ void __launch_bounds__(128, 1) mma_shapes_kernel(int iters, int mode, int N, long long *cycles) {
    extern __shared__ uint8_t raw[];
    Smem &s = *reinterpret_cast<Smem *>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (int)(sizeof(s.a) + sizeof(s.b)) / 2; i += blockDim.x) {
        uint32_t h = (i + blockIdx.x * 7919u) * 2654435761u;
        h ^= h >> 15;
        reinterpret_cast<uint16_t *>(s.a)[i] = static_cast<uint16_t>((h & 0x83FF) | 0x3000);   // fp16 in [0.125, 0.25), either sign
    }
    if (threadIdx.x == 0) {
        mbar_init(&s.done, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    if (warp == 2) tmem_alloc<512>(&s.tmem_base);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = (1u << 4) | ((static_cast<uint32_t>(N) >> 3) << 17) | ((128u >> 4) << 24) | ((mode == 1 || mode == 2) ? (1u << 16) : 0u);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const int st = i & 3, ks = (i >> 2) & 3;
            const uint64_t ad = make_sw128_kmajor_desc(smem_u32(s.a[st]) + ks * 32);
            const uint64_t bk = make_sw128_kmajor_desc(smem_u32(s.b[st]) + ks * 32);
            const uint64_t bmn = make_smem_desc(smem_u32(s.b[st]) + ks * 16 * 128, 128 * 128, 1024, kLayoutSw128);
            if (mode == 0) tc_mma_f16(tmem, ad, bk, idesc, i != 0);
            else if (mode == 1) tc_mma_f16(tmem, ad, bmn, idesc, i != 0);
            else if (mode == 2) mma_ts(tmem, tmem + 256 + ks * 8, bmn, idesc, i != 0);
            else mma_ts(tmem, tmem + 256 + ks * 8, bk, idesc, i != 0);
        }
        tc_commit(&s.done);
        mbar_wait(&s.done, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem);
    }
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4096;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long *cyc;
    cudaMalloc(&cyc, sms * sizeof(long long));
    const size_t smem = sizeof(Smem) + 1024;
    cudaFuncSetAttribute(mma_shapes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct Case { const char *name; int mode, N; } cases[] = {
        {"ss_kk_n64", 0, 64}, {"ss_kk_n128", 0, 128}, {"ss_kk_n256", 0, 256}, {"ss_kmn_n128", 1, 128}, {"ts_mn_n128", 2, 128}, {"ts_k_n128", 3, 128},
        {"ts_mn_n64", 2, 64}};
    printf("{\"iters\": %d, \"sms\": %d", iters, sms);
    for (const Case &c : cases) {
        for (int rep = 0; rep < 2; rep++) {
            mma_shapes_kernel<<<sms, 128, smem>>>(iters, c.mode, c.N, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf(", \"%s\": \"%s\"", c.name, cudaGetErrorString(e));
                break;
            }
            if (rep == 1) {
                std::vector<long long> h(sms);
                cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (long long v : h) mx = v > mx ? v : mx;
                printf(", \"%s_clk_per_mma\": %.1f", c.name, (double)mx / iters);
            }
        }
    }
    printf("}\n");
    return 0;
}
