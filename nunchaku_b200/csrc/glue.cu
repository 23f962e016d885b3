// Elementwise / row-reduce glue between the SVDQuant linears:
//   Silu / GELU::forward         src/activation.cpp:4-14, kernels/activation_kernels_impl.cuh:7-10,93-97
//   LayerNorm / RMSNorm::forward src/layernorm.cpp:14-24, kernels/layernorm_kernels.cu:6-58,
//                                kernels/layernorm_kernels_impl.cuh:13-22,46-164,292-320
//   add, mul_add_batch, split_mod, cast   src/kernels/misc_kernels.cu, misc_kernels_impl.cuh:13-90,183-203
//
// All of them are HBM-bound: every element is read once with 16-byte loads and written once with
// 16-byte stores; the two norms keep the row in registers between the statistics pass and the
// normalise pass (the reference stages it in shared memory), one CTA of 256 threads per row.
// Per-element arithmetic keeps the reference's rounding points (which ops happen in the 16-bit type,
// which in fp32); the row statistics are fp32 sums in a fixed order (deterministic, but a different
// order from the reference's block reduce -> tolerance, not bit equality, on the norms).
#include <type_traits>

#include "common.cuh"

namespace nb200 {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxSplit = 6;  // split_mod<2..6>  (misc_kernels.cu:319-323)

template <typename T, int N>
struct alignas(sizeof(T) * N) Vec {
    T v[N];
};

template <typename T>
__device__ __forceinline__ float to_f(T v) {
    if constexpr (std::is_same_v<T, float>)
        return v;
    else
        return HalfTraits<T>::to_float(v);
}
template <typename T>
__device__ __forceinline__ T from_f(float v) {
    if constexpr (std::is_same_v<T, float>)
        return v;
    else
        return HalfTraits<T>::from_float(v);
}
template <typename T>
__device__ __forceinline__ T t_mul(T a, T b) {
    if constexpr (std::is_same_v<T, float>)
        return a * b;
    else
        return __hmul(a, b);
}
template <typename T>
__device__ __forceinline__ T t_add(T a, T b) {
    if constexpr (std::is_same_v<T, float>)
        return a + b;
    else
        return __hadd(a, b);
}
template <typename T>
__device__ __forceinline__ T clamp_half(T v) {  // fp16 results saturate instead of overflowing to inf
    if constexpr (std::is_same_v<T, __half>) {
        v = __hmin(v, __float2half_rn(65504.f));
        v = __hmax(v, __float2half_rn(-65504.f));
    }
    return v;
}

// ---- activations ------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T silu_ref(T x) {  // activation_kernels_impl.cuh:7-10
    return from_f<T>(to_f(x) / (1.0f + expf(-to_f(x))));
}
template <typename T>
__device__ __forceinline__ T gelu_new_ref(T x) {  // activation_kernels_impl.cuh:93-97: mixed T / fp32 chain
    const float x3 = to_f(t_mul(t_mul(x, x), x));
    const T inner = t_add(x, from_f<T>(0.044715f * x3));
    const T arg = from_f<T>(0.79788456f * to_f(inner));
    const T t = from_f<T>(tanhf(to_f(arg)));
    return t_mul(t_mul(from_f<T>(0.5f), x), t_add(from_f<T>(1.0f), t));
}

template <typename T, int ACT>
__global__ void __launch_bounds__(kThreads) activation_kernel(const T *__restrict__ x, T *__restrict__ out, long long numel) {
    constexpr int V = 16 / sizeof(T);
    const long long nvec = numel / V;
    for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        Vec<T, V> a = reinterpret_cast<const Vec<T, V> *>(x)[i];
#pragma unroll
        for (int k = 0; k < V; k++) a.v[k] = ACT == 0 ? silu_ref(a.v[k]) : gelu_new_ref(a.v[k]);
        reinterpret_cast<Vec<T, V> *>(out)[i] = a;
    }
    if (blockIdx.x == 0 && threadIdx.x < numel - nvec * V) {  // ragged tail
        const long long i = nvec * V + threadIdx.x;
        out[i] = ACT == 0 ? silu_ref(x[i]) : gelu_new_ref(x[i]);
    }
}

// ---- add / mul_add_batch / split_mod / cast -------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) add_kernel(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ c,
                                                       long long numel) {
    constexpr int V = 16 / sizeof(T);
    const long long nvec = numel / V;
    for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        Vec<T, V> ra = reinterpret_cast<const Vec<T, V> *>(a)[i];
        const Vec<T, V> rb = reinterpret_cast<const Vec<T, V> *>(b)[i];
#pragma unroll
        for (int k = 0; k < V; k++) ra.v[k] = t_add(ra.v[k], rb.v[k]);
        reinterpret_cast<Vec<T, V> *>(c)[i] = ra;
    }
    if (blockIdx.x == 0 && threadIdx.x < numel - nvec * V) {
        const long long i = nvec * V + threadIdx.x;
        c[i] = t_add(a[i], b[i]);
    }
}

// x[b, i] = x[b, i] * (scale[b?, i % mod_scale] + shift) + bias[b?, i % mod_bias]   (misc_kernels_impl.cuh:25-68)
template <typename T, bool NO_SCALE>
__global__ void __launch_bounds__(kThreads)
mul_add_kernel(T *__restrict__ x, const T *__restrict__ scale, const T *__restrict__ bias, float scale_shift_f, long long numel,
               long long mod_scale, long long mod_bias, long long stride_x, long long stride_scale, long long stride_bias) {
    constexpr int V = 8;  // the reference's unroll: alignment contract of scale / bias is 8 elements
    const int b = blockIdx.y;
    const T shift = from_f<T>(scale_shift_f);
    const long long nvec = numel / V;
    for (long long j = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; j < nvec;
         j += static_cast<long long>(gridDim.x) * kThreads) {
        const long long i = j * V;
        Vec<T, V> rx = *reinterpret_cast<const Vec<T, V> *>(x + i + stride_x * b);
        const Vec<T, V> rb = *reinterpret_cast<const Vec<T, V> *>(bias + (i % mod_bias) + stride_bias * b);
        Vec<T, V> rs;
        if constexpr (!NO_SCALE) rs = *reinterpret_cast<const Vec<T, V> *>(scale + (i % mod_scale) + stride_scale * b);
#pragma unroll
        for (int k = 0; k < V; k++) {
            T t;
            if constexpr (NO_SCALE)
                t = t_add(rx.v[k], rb.v[k]);
            else
                t = t_add(t_mul(rx.v[k], t_add(rs.v[k], shift)), rb.v[k]);
            rx.v[k] = clamp_half(t);
        }
        *reinterpret_cast<Vec<T, V> *>(x + i + stride_x * b) = rx;
    }
}

struct SplitPtrs {
    void *p[kMaxSplit];
};

// input [..., C*N] interleaved -> N tensors [..., C]: out[k][i] = in[i*N + k]  (misc_kernels_impl.cuh:81-90).
// One thread gathers 8 consecutive outputs of every k from 8*N consecutive inputs (16-byte stores).
template <typename T, int N>
__global__ void __launch_bounds__(kThreads) split_mod_kernel(const T *__restrict__ in, SplitPtrs outs, long long per_out) {
    constexpr int V = 16 / sizeof(T);
    const long long nvec = per_out / V;
    for (long long j = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; j < nvec;
         j += static_cast<long long>(gridDim.x) * kThreads) {
        T buf[V * N];
#pragma unroll
        for (int c = 0; c < N; c++)
            *reinterpret_cast<Vec<T, V> *>(buf + c * V) = *reinterpret_cast<const Vec<T, V> *>(in + (j * N + c) * V);
#pragma unroll
        for (int k = 0; k < N; k++) {
            Vec<T, V> o;
#pragma unroll
            for (int e = 0; e < V; e++) o.v[e] = buf[e * N + k];
            reinterpret_cast<Vec<T, V> *>(outs.p[k])[j] = o;
        }
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * V + threadIdx.x; i < per_out; i += kThreads)
#pragma unroll
            for (int k = 0; k < N; k++) static_cast<T *>(outs.p[k])[i] = in[i * N + k];
    }
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kThreads) cast_kernel(const Tin *__restrict__ in, Tout *__restrict__ out, long long numel) {
    constexpr int V = 16 / (sizeof(Tin) > sizeof(Tout) ? sizeof(Tin) : sizeof(Tout));
    const long long nvec = numel / V;
    auto conv = [](Tin v) -> Tout {
        Tout o;
        if constexpr (std::is_same_v<Tin, Tout>)
            o = v;
        else
            o = from_f<Tout>(to_f(v));  // every pair goes through fp32 exactly (16-bit -> fp32 is exact)
        return clamp_half(o);
    };
    for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        const Vec<Tin, V> a = reinterpret_cast<const Vec<Tin, V> *>(in)[i];
        Vec<Tout, V> o;
#pragma unroll
        for (int k = 0; k < V; k++) o.v[k] = conv(a.v[k]);
        reinterpret_cast<Vec<Tout, V> *>(out)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < numel - nvec * V) {
        const long long i = nvec * V + threadIdx.x;
        out[i] = conv(in[i]);
    }
}

// ---- row norms -----------------------------------------------------------------------------------
// Fixed-order block reduction of up to two fp32 values: lanes by xor-shuffle, then warp 0 over the 8 partials.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float *sh /* [NV][8] */) {
#pragma unroll
    for (int k = 0; k < NV; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();  // sh may still be read from a previous call
    if (l == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) sh[k * 8 + w] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < kThreads / 32; i++) t += sh[k * 8 + i];
        v[k] = t;
    }
}

// LayerNorm (layernorm_general -> generalLayerNorm<T2, half, USE_DIFF_OF_SQUARES = true>):
//   mean = sum/H, var = sumsq/H - mean^2, y = T((x - mean) * rsqrtf(var + eps) [* gamma] [+ beta])   all fp32
// RMSNorm (rms_norm_kernel):  y = T(x * rsqrtf(sumsq/H + eps)) *_T weight
// CHUNKS 16-byte pieces per thread stay in registers; rows longer than CHUNKS*kThreads*8 re-read x (L2-hot).
template <typename T, bool RMS, int CHUNKS>
__global__ void __launch_bounds__(kThreads)
norm_kernel(const T *__restrict__ x, const T *__restrict__ gamma, const T *__restrict__ beta, T *__restrict__ out, int hidden,
            float eps) {
    constexpr int V = 8;
    __shared__ float sh[16];
    const long long row = blockIdx.x;
    const T *xr = x + row * hidden;
    T *orow = out + row * hidden;
    const int nvec = hidden / V;
    Vec<T, V> reg[CHUNKS];
    float acc[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < CHUNKS; c++) {
        const int j = c * kThreads + threadIdx.x;
        if (j < nvec) {
            reg[c] = reinterpret_cast<const Vec<T, V> *>(xr)[j];
#pragma unroll
            for (int k = 0; k < V; k++) {
                const float f = to_f(reg[c].v[k]);
                acc[0] += f;
                acc[1] = fmaf(f, f, acc[1]);
            }
        }
    }
    for (int j = CHUNKS * kThreads + threadIdx.x; j < nvec; j += kThreads) {  // long rows: statistics from a streamed pass
        const Vec<T, V> a = reinterpret_cast<const Vec<T, V> *>(xr)[j];
#pragma unroll
        for (int k = 0; k < V; k++) {
            const float f = to_f(a.v[k]);
            acc[0] += f;
            acc[1] = fmaf(f, f, acc[1]);
        }
    }
    block_sum<2>(acc, sh);
    float mean = 0.f, rstd;
    if constexpr (RMS) {
        rstd = rsqrtf(acc[1] / hidden + eps);
    } else {
        mean = acc[0] / hidden;
        rstd = rsqrtf(acc[1] / hidden - mean * mean + eps);
    }
    auto apply = [&](Vec<T, V> a, int j) {
        Vec<T, V> g, b;
        if (gamma != nullptr) g = reinterpret_cast<const Vec<T, V> *>(gamma)[j];
        if (!RMS && beta != nullptr) b = reinterpret_cast<const Vec<T, V> *>(beta)[j];
#pragma unroll
        for (int k = 0; k < V; k++) {
            if constexpr (RMS) {
                const T n = from_f<T>(to_f(a.v[k]) * rstd);
                a.v[k] = gamma != nullptr ? t_mul(n, g.v[k]) : n;
            } else {
                float r = (to_f(a.v[k]) - mean) * rstd;
                if (gamma != nullptr) r = r * to_f(g.v[k]);
                if (beta != nullptr) r = r + to_f(b.v[k]);
                a.v[k] = from_f<T>(r);
            }
        }
        reinterpret_cast<Vec<T, V> *>(orow)[j] = a;
    };
#pragma unroll
    for (int c = 0; c < CHUNKS; c++) {
        const int j = c * kThreads + threadIdx.x;
        if (j < nvec) apply(reg[c], j);
    }
    for (int j = CHUNKS * kThreads + threadIdx.x; j < nvec; j += kThreads) apply(reinterpret_cast<const Vec<T, V> *>(xr)[j], j);
}

// Rows up to 4096 elements (every norm on the FLUX / SANA path): ONE WARP per row, the row lives in registers,
// statistics by xor-shuffle only -- no shared memory, no __syncthreads, ~6 KB of loads in flight per warp.
template <typename T, bool RMS, int MAXC>
__global__ void __launch_bounds__(kThreads)
norm_warp_kernel(const T *__restrict__ x, const T *__restrict__ gamma, const T *__restrict__ beta, T *__restrict__ out,
                 long long rows, int hidden, float eps, const T *__restrict__ mod_scale = nullptr, const T *__restrict__ mod_shift = nullptr,
                 float mod_scale_shift = 0.f) {
    constexpr int V = 8;
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T *xr = x + row * hidden;
    T *orow = out + row * hidden;
    const int nvec = hidden / V;
    Vec<T, V> reg[MAXC];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; c++) {
        const int j = c * 32 + lane;
        if (j < nvec) reg[c] = reinterpret_cast<const Vec<T, V> *>(xr)[j];
    }
#pragma unroll
    for (int c = 0; c < MAXC; c++) {
        const int j = c * 32 + lane;
        if (j < nvec) {
#pragma unroll
            for (int k = 0; k < V; k++) {
                const float f = to_f(reg[c].v[k]);
                s0 += f;
                s1 = fmaf(f, f, s1);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    float mean = 0.f, rstd;
    if constexpr (RMS) {
        rstd = rsqrtf(s1 / hidden + eps);
    } else {
        mean = s0 / hidden;
        rstd = rsqrtf(s1 / hidden - mean * mean + eps);
    }
#pragma unroll
    for (int c = 0; c < MAXC; c++) {
        const int j = c * 32 + lane;
        if (j < nvec) {
            Vec<T, V> a = reg[c], g, b;
            if (gamma != nullptr) g = reinterpret_cast<const Vec<T, V> *>(gamma)[j];
            if (!RMS && beta != nullptr) b = reinterpret_cast<const Vec<T, V> *>(beta)[j];
#pragma unroll
            for (int k = 0; k < V; k++) {
                if constexpr (RMS) {
                    const T n = from_f<T>(to_f(a.v[k]) * rstd);
                    a.v[k] = gamma != nullptr ? t_mul(n, g.v[k]) : n;
                } else {
                    float r = (to_f(a.v[k]) - mean) * rstd;
                    if (gamma != nullptr) r = r * to_f(g.v[k]);
                    if (beta != nullptr) r = r + to_f(b.v[k]);
                    a.v[k] = from_f<T>(r);
                }
            }
            if (mod_shift != nullptr) {   // AdaLN: the mul_add kernel's arithmetic on the rounded norm (nb200_layernorm_mod): y * (scale + c) + shift in T
                const Vec<T, V> ms = reinterpret_cast<const Vec<T, V> *>(mod_scale)[j], mb = reinterpret_cast<const Vec<T, V> *>(mod_shift)[j];
                const T c = from_f<T>(mod_scale_shift);
#pragma unroll
                for (int k = 0; k < V; k++) a.v[k] = clamp_half(t_add(t_mul(a.v[k], t_add(ms.v[k], c)), mb.v[k]));
            }
            reinterpret_cast<Vec<T, V> *>(orow)[j] = a;
        }
    }
}

int grid_for(long long nvec) {
    long long g = (nvec + kThreads - 1) / kThreads;
    const long long cap = 148LL * 16;  // grid-stride beyond 16 resident CTAs per SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename F>
int dispatch_dtype(int dtype, F &&f) {
    switch (dtype) {
    case NB200_FP16: return f(__half{});
    case NB200_BF16: return f(__nv_bfloat16{});
    case NB200_FP32: return f(float{});
    default: return fail(NB200_ERR_INVALID_ARGUMENT, "dtype must be NB200_FP16, NB200_BF16 or NB200_FP32");
    }
}

}  // namespace
}  // namespace nb200

using namespace nb200;

#pragma GCC visibility push(default)
extern "C" {

int nb200_activation(int kind, int dtype, const void *x, void *out, long long numel, void *stream_) {
    NB200_REQUIRE(x != nullptr && out != nullptr, "activation: null tensor");
    NB200_REQUIRE(kind == NB200_ACT_SILU || kind == NB200_ACT_GELU, "activation: kind must be NB200_ACT_SILU or NB200_ACT_GELU");
    NB200_REQUIRE(numel >= 0, "activation: negative size");
    NB200_REQUIRE(aligned16(x) && aligned16(out), "activation: tensors must be 16-byte aligned");
    if (numel == 0) return NB200_OK;
    auto stream = static_cast<cudaStream_t>(stream_);
    return dispatch_dtype(dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        const int grid = grid_for(numel / (16 / sizeof(T)));
        if (kind == NB200_ACT_SILU)
            activation_kernel<T, 0><<<grid, kThreads, 0, stream>>>(static_cast<const T *>(x), static_cast<T *>(out), numel);
        else
            activation_kernel<T, 1><<<grid, kThreads, 0, stream>>>(static_cast<const T *>(x), static_cast<T *>(out), numel);
        NB200_CUDA_CHECK(cudaGetLastError());
        count_launch();
        return NB200_OK;
    });
}

int nb200_add(int dtype, const void *a, const void *b, void *out, long long numel, void *stream_) {
    NB200_REQUIRE(a != nullptr && b != nullptr && out != nullptr, "add: null tensor");
    NB200_REQUIRE(numel >= 0, "add: negative size");
    NB200_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out), "add: tensors must be 16-byte aligned");
    if (numel == 0) return NB200_OK;
    auto stream = static_cast<cudaStream_t>(stream_);
    return dispatch_dtype(dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        add_kernel<T><<<grid_for(numel / (16 / sizeof(T))), kThreads, 0, stream>>>(static_cast<const T *>(a), static_cast<const T *>(b),
                                                                                  static_cast<T *>(out), numel);
        NB200_CUDA_CHECK(cudaGetLastError());
        count_launch();
        return NB200_OK;
    });
}

int nb200_mul_add_batch(int dtype, void *x, const void *scale, const void *bias, float scale_shift, int batch, long long numel,
                        long long numel_scale, long long numel_bias, long long stride_x, long long stride_scale,
                        long long stride_bias, void *stream_) {
    NB200_REQUIRE(x != nullptr && bias != nullptr, "mul_add_batch: x and bias are required");
    NB200_REQUIRE(batch >= 1 && numel >= 0, "mul_add_batch: bad sizes");
    if (scale == nullptr) numel_scale = 1;
    NB200_REQUIRE(numel_bias > 0 && numel % numel_bias == 0, "mul_add_batch: numel must be a multiple of bias.numel()");
    NB200_REQUIRE(scale == nullptr || (numel_scale > 0 && numel % numel_scale == 0),
                  "mul_add_batch: numel must be a multiple of scale.numel()");
    // misc_kernels.cu:96-102: 8-element vectors on every operand
    NB200_REQUIRE(numel % 8 == 0 && numel_bias % 8 == 0 && (scale == nullptr || numel_scale % 8 == 0),
                  "mul_add_batch: sizes must be multiples of 8");
    NB200_REQUIRE(stride_x % 8 == 0 && stride_scale % 8 == 0 && stride_bias % 8 == 0, "mul_add_batch: batch strides must be multiples of 8");
    if (numel == 0) return NB200_OK;
    auto stream = static_cast<cudaStream_t>(stream_);
    return dispatch_dtype(dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        const uintptr_t al = sizeof(T) * 8 - 1;
        NB200_REQUIRE((reinterpret_cast<uintptr_t>(x) & al) == 0 && (reinterpret_cast<uintptr_t>(bias) & al) == 0 &&
                          (reinterpret_cast<uintptr_t>(scale) & al) == 0,
                      "mul_add_batch: pointers must be aligned to 8 elements");
        dim3 grid(grid_for(numel / 8), batch);
        if (scale != nullptr)
            mul_add_kernel<T, false><<<grid, kThreads, 0, stream>>>(static_cast<T *>(x), static_cast<const T *>(scale),
                                                                    static_cast<const T *>(bias), scale_shift, numel, numel_scale,
                                                                    numel_bias, stride_x, stride_scale, stride_bias);
        else
            mul_add_kernel<T, true><<<grid, kThreads, 0, stream>>>(static_cast<T *>(x), nullptr, static_cast<const T *>(bias),
                                                                   scale_shift, numel, 1, numel_bias, stride_x, 0, stride_bias);
        NB200_CUDA_CHECK(cudaGetLastError());
        count_launch();
        return NB200_OK;
    });
}

int nb200_split_mod(int dtype, const void *input, void *const *outs, int n, long long numel, void *stream_) {
    NB200_REQUIRE(input != nullptr && outs != nullptr, "split_mod: null tensor");
    NB200_REQUIRE(n >= 2 && n <= kMaxSplit, "split_mod: 2..6 outputs (misc_kernels.cu:319-323)");
    NB200_REQUIRE(numel >= 0 && numel % n == 0, "split_mod: numel must be a multiple of the number of outputs");
    SplitPtrs p{};
    for (int k = 0; k < n; k++) {
        NB200_REQUIRE(outs[k] != nullptr && aligned16(outs[k]), "split_mod: outputs must be non-null and 16-byte aligned");
        p.p[k] = outs[k];
    }
    NB200_REQUIRE(aligned16(input), "split_mod: input must be 16-byte aligned");
    if (numel == 0) return NB200_OK;
    auto stream = static_cast<cudaStream_t>(stream_);
    return dispatch_dtype(dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        const long long per_out = numel / n;
        const int grid = grid_for(per_out / (16 / sizeof(T)));
        const T *in = static_cast<const T *>(input);
        switch (n) {
        case 2: split_mod_kernel<T, 2><<<grid, kThreads, 0, stream>>>(in, p, per_out); break;
        case 3: split_mod_kernel<T, 3><<<grid, kThreads, 0, stream>>>(in, p, per_out); break;
        case 4: split_mod_kernel<T, 4><<<grid, kThreads, 0, stream>>>(in, p, per_out); break;
        case 5: split_mod_kernel<T, 5><<<grid, kThreads, 0, stream>>>(in, p, per_out); break;
        default: split_mod_kernel<T, 6><<<grid, kThreads, 0, stream>>>(in, p, per_out); break;
        }
        NB200_CUDA_CHECK(cudaGetLastError());
        count_launch();
        return NB200_OK;
    });
}

int nb200_cast(int dtype_in, const void *input, int dtype_out, void *output, long long numel, void *stream_) {
    NB200_REQUIRE(input != nullptr && output != nullptr, "cast: null tensor");
    NB200_REQUIRE(numel >= 0, "cast: negative size");
    NB200_REQUIRE(aligned16(input) && aligned16(output), "cast: tensors must be 16-byte aligned");
    if (numel == 0) return NB200_OK;
    auto stream = static_cast<cudaStream_t>(stream_);
    return dispatch_dtype(dtype_in, [&](auto tin) -> int {
        using Tin = decltype(tin);
        return dispatch_dtype(dtype_out, [&](auto tout) -> int {
            using Tout = decltype(tout);
            if (input == output) NB200_REQUIRE(sizeof(Tin) == sizeof(Tout), "cast: in-place needs equal element sizes (misc_kernels.cu:263-265)");
            constexpr int V = 16 / (sizeof(Tin) > sizeof(Tout) ? sizeof(Tin) : sizeof(Tout));
            cast_kernel<Tin, Tout><<<grid_for(numel / V), kThreads, 0, stream>>>(static_cast<const Tin *>(input),
                                                                               static_cast<Tout *>(output), numel);
            NB200_CUDA_CHECK(cudaGetLastError());
            count_launch();
            return NB200_OK;
        });
    });
}

static int launch_norm(bool rms, int dtype, const void *x, const void *weight, const void *bias, void *out, long long rows,
                       int hidden, float eps, cudaStream_t stream) {
    NB200_REQUIRE(x != nullptr && out != nullptr, "norm: null tensor");
    NB200_REQUIRE(rows >= 0 && hidden > 0 && hidden % 8 == 0, "norm: hidden size must be a positive multiple of 8");
    NB200_REQUIRE(rows <= 0x7fffffffLL, "norm: too many rows");
    NB200_REQUIRE(!rms || weight != nullptr, "rms_norm: weight is required (layernorm_kernels_impl.cuh:317)");
    NB200_REQUIRE(aligned16(x) && aligned16(out) && aligned16(weight) && aligned16(bias), "norm: tensors must be 16-byte aligned");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "norm: 16-bit dtypes only");
    if (rows == 0) return NB200_OK;
    return dispatch_dtype(dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        if constexpr (std::is_same_v<T, float>) {
            return fail(NB200_ERR_INVALID_ARGUMENT, "norm: 16-bit dtypes only");
        } else {
            const T *xp = static_cast<const T *>(x), *g = static_cast<const T *>(weight), *b = static_cast<const T *>(bias);
            T *o = static_cast<T *>(out);
            if (hidden <= 4096) {
                const int wgrid = static_cast<int>((rows + kThreads / 32 - 1) / (kThreads / 32));
                const int wc = (hidden / 8 + 31) / 32;
#define NB200_NORMW(RMS, C) norm_warp_kernel<T, RMS, C><<<wgrid, kThreads, 0, stream>>>(xp, g, b, o, rows, hidden, eps)
                if (rms) {
                    if (wc <= 1) NB200_NORMW(true, 1);
                    else if (wc <= 4) NB200_NORMW(true, 4);
                    else if (wc <= 8) NB200_NORMW(true, 8);
                    else if (wc <= 12) NB200_NORMW(true, 12);
                    else NB200_NORMW(true, 16);
                } else {
                    if (wc <= 1) NB200_NORMW(false, 1);
                    else if (wc <= 4) NB200_NORMW(false, 4);
                    else if (wc <= 8) NB200_NORMW(false, 8);
                    else if (wc <= 12) NB200_NORMW(false, 12);
                    else NB200_NORMW(false, 16);
                }
#undef NB200_NORMW
                NB200_CUDA_CHECK(cudaGetLastError());
                count_launch();
                return NB200_OK;
            }
            const int grid = static_cast<int>(rows);
            const int chunks = (hidden / 8 + kThreads - 1) / kThreads;
#define NB200_NORM(RMS, C) norm_kernel<T, RMS, C><<<grid, kThreads, 0, stream>>>(xp, g, b, o, hidden, eps)
            if (rms) {
                if (chunks <= 1) NB200_NORM(true, 1);
                else if (chunks <= 2) NB200_NORM(true, 2);
                else NB200_NORM(true, 4);
            } else {
                if (chunks <= 1) NB200_NORM(false, 1);
                else if (chunks <= 2) NB200_NORM(false, 2);
                else NB200_NORM(false, 4);
            }
#undef NB200_NORM
            NB200_CUDA_CHECK(cudaGetLastError());
            count_launch();
            return NB200_OK;
        }
    });
}

int nb200_layernorm(int dtype, const void *x, const void *weight, const void *bias, void *out, long long rows, int hidden,
                    float eps, void *stream) {
    return launch_norm(false, dtype, x, weight, bias, out, rows, hidden, eps, static_cast<cudaStream_t>(stream));
}

// LayerNorm followed by the AdaLN modulation  y * (mod_scale[c] + scale_shift) + mod_shift[c]  in ONE pass over the rows (SURVEY section 8f row N2:
// the reference runs layernorm + mul_add_batch, src/FluxModel.cpp AdaLayerNormZero::forward); bit-identical to nb200_layernorm followed by
// nb200_mul_add_batch (same rounding points), one read + one write of the activations instead of two.
int nb200_layernorm_mod(int dtype, const void *x, const void *weight, const void *bias, const void *mod_scale, const void *mod_shift, float scale_shift,
                        void *out, long long rows, int hidden, float eps, void *stream_) {
    reset_launch_count();
    NB200_REQUIRE(x && out && mod_scale && mod_shift, "NULL tensor");
    NB200_REQUIRE(rows >= 0 && hidden > 0 && hidden % 8 == 0, "hidden must be a positive multiple of 8");
    NB200_REQUIRE(dtype == NB200_FP16 || dtype == NB200_BF16, "dtype must be fp16 or bf16");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows == 0) return NB200_OK;
    if (hidden > 4096) {   // rows too long for the one-warp-per-row kernel: the two launches
        if (int rc = launch_norm(false, dtype, x, weight, bias, out, rows, hidden, eps, stream)) return rc;
        return nb200_mul_add_batch(dtype, out, mod_scale, mod_shift, scale_shift, 1, rows * hidden, hidden, hidden, 0, 0, 0, stream_);
    }
    const int wgrid = static_cast<int>((rows + kThreads / 32 - 1) / (kThreads / 32));
    const int wc = (hidden / 8 + 31) / 32;
    auto go = [&](auto tag) {
        using T = decltype(tag);
        const T *xp = static_cast<const T *>(x), *g = static_cast<const T *>(weight), *b = static_cast<const T *>(bias);
        const T *ms = static_cast<const T *>(mod_scale), *mb = static_cast<const T *>(mod_shift);
        T *o = static_cast<T *>(out);
#define NB200_NORMW(C) norm_warp_kernel<T, false, C><<<wgrid, kThreads, 0, stream>>>(xp, g, b, o, rows, hidden, eps, ms, mb, scale_shift)
        if (wc <= 1) NB200_NORMW(1);
        else if (wc <= 4) NB200_NORMW(4);
        else if (wc <= 8) NB200_NORMW(8);
        else if (wc <= 12) NB200_NORMW(12);
        else NB200_NORMW(16);
#undef NB200_NORMW
    };
    if (dtype == NB200_BF16) go(__nv_bfloat16{});
    else go(__half{});
    NB200_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return NB200_OK;
}

int nb200_rms_norm(int dtype, const void *x, const void *weight, void *out, long long rows, int hidden, float eps, void *stream) {
    return launch_norm(true, dtype, x, weight, nullptr, out, rows, hidden, eps, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
#pragma GCC visibility pop
