// Micro-benchmark: issue rate of the ops the INT4 converter uses (per SM, 32 warps resident).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench/alu_rates.cu -o gpurun_out/alu_rates && ./gpurun_out/alu_rates
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int KIND>
__global__ void __launch_bounds__(1024, 1) k(unsigned *out, unsigned seed, long long *clk) {
    unsigned v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = seed + threadIdx.x * 17 + i;
    const unsigned s = seed | 0x3f803f80u, off = 0x43084308u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (KIND == 0) {  // HFMA2.BF16
                __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162 *>(&v[i]);
                a = __hfma2(a, *reinterpret_cast<const __nv_bfloat162 *>(&s), *reinterpret_cast<const __nv_bfloat162 *>(&off));
                v[i] = *reinterpret_cast<unsigned *>(&a);
            } else if (KIND == 1) {  // HFMA2 fp16
                __half2 a = *reinterpret_cast<__half2 *>(&v[i]);
                a = __hfma2(a, *reinterpret_cast<const __half2 *>(&s), *reinterpret_cast<const __half2 *>(&off));
                v[i] = *reinterpret_cast<unsigned *>(&a);
            } else if (KIND == 2) {  // HSUB2 + HMUL2 bf16 (the converter's pair)
                __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162 *>(&v[i]);
                a = __hmul2(__hsub2(a, *reinterpret_cast<const __nv_bfloat162 *>(&off)), *reinterpret_cast<const __nv_bfloat162 *>(&s));
                v[i] = *reinterpret_cast<unsigned *>(&a);
            } else if (KIND == 3) {  // FFMA
                float a = __uint_as_float(v[i]);
                a = fmaf(a, __uint_as_float(s), __uint_as_float(off));
                v[i] = __float_as_uint(a);
            } else if (KIND == 4) {  // LOP3 (and-or)
                v[i] = ((v[i] >> 4) & 0x000F000Fu) | 0x43004300u | (v[i] << 28);
            } else if (KIND == 5) {  // HSUB2 + HMUL2 fp16
                __half2 a = *reinterpret_cast<__half2 *>(&v[i]);
                a = __hmul2(__hsub2(a, *reinterpret_cast<const __half2 *>(&off)), *reinterpret_cast<const __half2 *>(&s));
                v[i] = *reinterpret_cast<unsigned *>(&a);
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char *name, int ops_per_inner) {
    unsigned *out;
    long long *clk;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&clk, 148 * 8);
    k<KIND><<<148, 1024>>>(out, 12345u, clk);
    k<KIND><<<148, 1024>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += h[i];
    avg /= 148;
    const double warp_instr = 32.0 * ITERS * ILP * ops_per_inner;  // per SM
    printf("%-28s %10.0f clk  -> %.3f warp-instr/clk/SM  (%.2f clk per warp-instr per SMSP)\n", name, avg, warp_instr / avg,
           avg / (warp_instr / 4));
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    run<0>("HFMA2.BF16", 1);
    run<1>("HFMA2.F16", 1);
    run<2>("HSUB2+HMUL2 bf16", 2);
    run<5>("HSUB2+HMUL2 f16", 2);
    run<3>("FFMA", 1);
    run<4>("SHF+LOP3 chain", 3);
    return 0;
}
