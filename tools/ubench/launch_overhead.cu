// What does one launch cost on a B200 when the kernel does nothing?  CUDA events around single launches of an empty kernel, as a function of
//   * its dynamic shared memory (0 / 100 / 217 KB: the quantizer and the GEMMs ask for > 200 KB, which changes the SM's L1 / shared carve-out),
//   * what ran before it (nothing / a 0-smem elementwise kernel / the same big-smem kernel),
//   * cluster launch (2 CTAs) and the programmatic-dependent-launch attribute,
//   * a body that only initialises ~80 mbarriers and allocates / frees 512 TMEM columns (the fixed part of the GEMM's prologue).
// r02 question: the fused GEMM spends ~7 us per launch outside any CTA's lifetime and the quantizer has a 12 us floor with its loads and math
// switched off (tools/quant_ablate.py) -- launch machinery or kernel prologue?
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I nunchaku_b200/csrc tools/ubench/launch_overhead.cu -o tools/ubench/_bin/launch_overhead
#include <algorithm>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace nb200::ptx;

__global__ void empty_kernel(int *sink) {
    extern __shared__ uint8_t smem[];
    if (sink != nullptr && threadIdx.x == 1023) sink[0] = smem[0];
}
__global__ void __launch_bounds__(384, 1) prologue_kernel(int *sink, int nbar, int tmem) {
    extern __shared__ uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    __shared__ uint32_t tmem_base;
    griddep_launch_dependents();
    if (threadIdx.x == 0) {
        for (int i = 0; i < nbar; i++) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    if (tmem && threadIdx.x < 32) tmem_alloc<512>(&tmem_base);
    __syncthreads();
    griddep_wait();
    if (tmem && threadIdx.x < 32) tmem_dealloc<512>(tmem_base);
    if (sink != nullptr && threadIdx.x == 1023) sink[0] = smem[0];
}
__global__ void elementwise_kernel(float *x, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += 1.f;
}

struct Case {
    const char *name;
    int smem_kb, before, cluster, pdl, body;   // before: 0 nothing, 1 elementwise, 2 same kernel;  body: 0 empty, 1 barriers, 2 barriers + TMEM
};

int main() {
    float *x;
    cudaMalloc(&x, 64 << 20);
    cudaFuncSetAttribute(empty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(prologue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(empty_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const Case cases[] = {
        {"empty, 0 KB smem, after nothing", 0, 0, 1, 0, 0},
        {"empty, 0 KB smem, after elementwise", 0, 1, 1, 0, 0},
        {"empty, 100 KB smem, after elementwise", 100, 1, 1, 0, 0},
        {"empty, 217 KB smem, after elementwise", 217, 1, 1, 0, 0},
        {"empty, 217 KB smem, after itself", 217, 2, 1, 0, 0},
        {"empty, 217 KB smem, after elementwise, PDL attr", 217, 1, 1, 1, 0},
        {"empty, 217 KB smem, cluster of 2, after elementwise", 217, 1, 2, 0, 0},
        {"empty, 217 KB smem, cluster of 2, after itself", 217, 2, 2, 0, 0},
        {"80 mbarrier inits, 217 KB, after elementwise", 217, 1, 1, 0, 1},
        {"80 mbarrier inits + TMEM alloc/free, 217 KB, after elementwise", 217, 1, 1, 0, 2},
        {"80 mbarrier inits + TMEM alloc/free, 217 KB, after itself, PDL", 217, 2, 1, 1, 2},
    };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("{\"what\": \"event-to-event time of one launch of a do-nothing kernel, 148 CTAs x 384 threads\", \"rows\": [\n");
    bool first = true;
    for (const Case &c : cases) {
        std::vector<float> ts;
        for (int it = 0; it < 30; it++) {
            cudaLaunchConfig_t cfg{};
            cudaLaunchAttribute attr[2];
            unsigned n = 0;
            cfg.gridDim = dim3(148);
            cfg.blockDim = dim3(384);
            cfg.dynamicSmemBytes = size_t(c.smem_kb) * 1024;
            if (c.cluster > 1) {
                attr[n].id = cudaLaunchAttributeClusterDimension;
                attr[n].val.clusterDim.x = c.cluster;
                attr[n].val.clusterDim.y = 1;
                attr[n].val.clusterDim.z = 1;
                n++;
            }
            if (c.pdl) {
                attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[n].val.programmaticStreamSerializationAllowed = 1;
                n++;
            }
            cfg.attrs = attr;
            cfg.numAttrs = n;
            auto launch = [&]() {
                if (c.body == 0) cudaLaunchKernelEx(&cfg, empty_kernel, (int *)nullptr);
                else cudaLaunchKernelEx(&cfg, prologue_kernel, (int *)nullptr, 80, c.body == 2 ? 1 : 0);
            };
            // a long elementwise kernel first so that the CPU has enqueued everything before the GPU gets here
            elementwise_kernel<<<(16 << 20) / 256, 256>>>(x, 16 << 20);
            elementwise_kernel<<<(16 << 20) / 256, 256>>>(x, 16 << 20);
            if (c.before == 2) launch();
            if (c.before == 1) elementwise_kernel<<<148, 256>>>(x, 148 * 256);
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 5) ts.push_back(ms * 1e3f);
        }
        std::sort(ts.begin(), ts.end());
        printf("%s  {\"case\": \"%s\", \"median_us\": %.2f, \"min_us\": %.2f}", first ? "" : ",\n", c.name, ts[ts.size() / 2], ts[0]);
        first = false;
        if (cudaGetLastError() != cudaSuccess) fprintf(stderr, "CUDA error in case %s\n", c.name);
    }
    printf("\n]}\n");
    return 0;
}
