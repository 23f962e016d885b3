/* TEST INFRASTRUCTURE (oracle) -- plain-C restatement of oracle/svdq.py's reference-emulating path
 * (svdq_linear_forward(mode="ref", act="none")): activation quantise + low-rank down, W4A4 main loop with the
 * reference's rounding chain, bias / per-channel scale, low-rank up.  Same citations as the Python file:
 *   quantize_w4a4_fuse_lora_kernel      src/kernels/zgemm/gemm_w4a4.cuh:1097-1184
 *   quantize_w4a4_from_fpsum_warp       gemm_w4a4.cuh:429-523        (INT4)
 *   quantize_w4a4_fp4_from_fpsum_warp   gemm_w4a4.cuh:85-187         (NVFP4)
 *   gemm_w4a4_block + apply_scales      gemm_w4a4.cuh:831-928, gemm_base.cuh:368-409
 *   gemm_w4a4_fp4_block                 gemm_w4a4.cuh:273-356
 *   EpilogueBias, Lora::EpilogueLoraUp  gemm_base.cuh:710-781, lora.cuh:110-241
 * PARITY UNPINNED for the same reason as oracle/svdq.py (no CPU implementation / golden vectors in the reference);
 * tests/test_oracle_c.py pins this file against the Python restatement instead.
 *
 * It exists so that the CPU arm of bench.py (`cpu_baseline`, `--impl reference`) runs at the speed of compiled,
 * multi-threaded code rather than of a Python loop.  Only tests/, __graft_entry__ and bench.py may load it.
 *
 * All tensors are float32 arrays whose values are exactly representable in the 16-bit type hT (bf16 or fp16):
 *   x [M,K], smooth [K], lora_down [R,K], wscales (INT4: hT [N,K/64]; NVFP4: decoded ue4m3 [N,K/16]), bias [N],
 *   wcscales [N] or NULL, lora_up [N,R]; qw int8 [N,K] (INT4 -8..7 | e2m1 codes 0..15); out [M,N].
 * gcc -O2 -fopenmp -shared -fPIC oracle/svdq_ref.c -o oracle/_build/libsvdq_ref.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* float -> hT -> float, round to nearest even */
static inline float f32_to_ht(float f, int bf16) {
    if (!bf16) return (float)(_Float16)f;
    uint32_t u = f2u(f);
    if ((u & 0x7F800000u) == 0x7F800000u) return u2f(u & 0xFFFF0000u | ((u & 0xFFFFu) ? 0x00400000u : 0u)); /* inf / nan */
    u += 0x7FFFu + ((u >> 16) & 1u);
    return u2f(u & 0xFFFF0000u);
}

/* double -> hT with ONE rounding (what a fused hardware op does): fp16 via the correctly rounded double->_Float16
 * conversion, bf16 via float with round-to-odd followed by RNE (the extra 16 bits make the double rounding exact) */
static inline float f64_to_ht(double v, int bf16) {
    if (!bf16) return (float)(_Float16)v;
    float f = (float)v;
    double back = (double)f;
    if (isfinite(v) && isfinite(back) && back != v) {
        if (fabs(back) > fabs(v)) f = nextafterf(f, 0.0f);
        f = u2f(f2u(f) | 1u);
    }
    return f32_to_ht(f, 1);
}

static const double kE2M1[8] = {0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0};
static inline double e2m1_decode(int c) { double m = kE2M1[c & 7]; return (c & 8) ? -m : m; }

/* cvt.rn.satfinite.e2m1x2.f32 (gemm_utils.cuh:239-245) */
static inline int e2m1_encode(float v) {
    if (isnan(v)) return 7;
    double a = fabs((double)v);
    int code = 0;
    if (a > 0.25) code = 1;
    if (a >= 0.75) code = 2;
    if (a > 1.25) code = 3;
    if (a >= 1.75) code = 4;
    if (a > 2.5) code = 5;
    if (a >= 3.5) code = 6;
    if (a > 5.0) code = 7;
    return code | (signbit(v) ? 8 : 0);
}

/* cvt.rn.satfinite.e4m3x2.f32 for 0 <= v <= 448: returns the DECODED value of the ue4m3 code */
static inline float e4m3_round(float v) {
    if (!(v > 0.f)) return 0.f;
    if (v >= 448.f) return 448.f;
    int e;
    (void)frexpf(v, &e);          /* v = m * 2^e, m in [0.5, 1) */
    int ex = e - 1;               /* v in [2^ex, 2^(ex+1)) */
    if (ex < -6) ex = -6;         /* subnormals share the exponent of the smallest normal: step 2^-9 */
    float step = ldexpf(1.f, ex - 3);
    float q = nearbyintf(v / step) * step;   /* round to nearest even (default rounding mode) */
    return q > 448.f ? 448.f : q;
}

void svdq_linear_ref(int M, int K, int N, int R, int fp4, int bf16, const float *x, const float *smooth, const float *lora_down,
                     const int8_t *qw, const float *wscales, const float *bias, const float *wcscales, float alpha, const float *lora_up,
                     float *out) {
    const int Mp = (M + 255) / 256 * 256;
    const int G = fp4 ? K / 16 : K / 64, GS = fp4 ? 16 : 64;
    int8_t *qa = (int8_t *)malloc((size_t)Mp * K);
    float *as = (float *)malloc((size_t)Mp * G * sizeof(float));       /* [Mp][G] row-major */
    float *la = (float *)calloc((size_t)Mp * (R > 0 ? R : 1), sizeof(float));

#pragma omp parallel for schedule(static)
    for (int m = 0; m < Mp; m++) {
        /* rows >= M are zero (load_act_to_fpsum pads, gemm_base.cuh:592-646) */
        float *xs = (float *)malloc((size_t)K * sizeof(float));
        for (int k = 0; k < K; k++) {
            const float xv = m < M ? x[(size_t)m * K + k] : 0.f;
            xs[k] = smooth ? f32_to_ht(xv / smooth[k], bf16) : xv;   /* h2div: fp32 divide, rounded to hT */
        }
        /* EpilogueLoraDown on the UN-smoothed row, fp32 result of an (order-free) exact-ish sum: fp64 here */
        for (int r = 0; r < R; r++) {
            double s = 0.0;
            if (m < M)
                for (int k = 0; k < K; k++) s += (double)x[(size_t)m * K + k] * (double)lora_down[(size_t)r * K + k];
            la[(size_t)m * R + r] = (float)s;
        }
        for (int g = 0; g < G; g++) {
            float amax = 0.f;
            for (int i = 0; i < GS; i++) amax = fmaxf(amax, fabsf(xs[g * GS + i]));
            if (!fp4) {
                const float s32 = amax * (float)(1.0 / 7.0);               /* :485-486 fp32 multiply */
                as[(size_t)m * G + g] = f32_to_ht(s32, bf16);                /* stored scale */
                const float rs = (float)(1.0 / (double)s32);                 /* rcp.approx.ftz, exact here */
                for (int i = 0; i < GS; i++) {
                    float p = nearbyintf(xs[g * GS + i] * rs);               /* cvt.rni; NaN -> 0 */
                    if (isnan(p)) p = 0.f;
                    if (p < -8.f) p = -8.f;
                    if (p > 7.f) p = 7.f;
                    qa[(size_t)m * K + g * GS + i] = (int8_t)p;
                }
            } else {
                const float s32 = fminf(amax * (float)(1.0 / 6.0), 448.0f); /* :133-134 */
                as[(size_t)m * G + g] = e4m3_round(s32);                     /* :141-142, decoded */
                const float rs = (float)(1.0 / (double)s32);                 /* :137-138 unrounded scale */
                for (int i = 0; i < GS; i++) qa[(size_t)m * K + g * GS + i] = (int8_t)e2m1_encode(xs[g * GS + i] * rs);
            }
        }
        free(xs);
    }

    /* NVFP4: decode the activation codes once (values are exact in fp32); weight rows are decoded where they are used */
    float *adec = NULL;
    if (fp4) {
        adec = (float *)malloc((size_t)M * K * sizeof(float));
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)M * K; i++) adec[i] = (float)e2m1_decode(qa[i]);
    }

    /* one weight row per iteration, all M activation rows inside: the row of W stays in L1/L2 and the activations (a few MB)
     * in the shared cache, instead of streaming the whole weight matrix once per activation row */
#pragma omp parallel
    {
        float *wrow = fp4 ? (float *)malloc((size_t)K * sizeof(float)) : NULL;
#pragma omp for schedule(static)
        for (int n = 0; n < N; n++) {
            if (fp4)
                for (int k = 0; k < K; k++) wrow[k] = (float)e2m1_decode(qw[(size_t)n * K + k]);
            for (int m = 0; m < M; m++) {
                double acc;
                if (!fp4) {
                    /* acc_hT = fma_hT(hT(int32 p), mul_hT(as, ws), acc_hT) per 64-wide group */
                    float a = 0.f;
                    for (int g = 0; g < G; g++) {
                        int32_t p = 0;
                        const int8_t *pa = qa + (size_t)m * K + g * 64, *pw = qw + (size_t)n * K + g * 64;
                        for (int i = 0; i < 64; i++) p += (int32_t)pa[i] * (int32_t)pw[i];
                        const float ph = f32_to_ht((float)p, bf16);
                        const float sc = f64_to_ht((double)as[(size_t)m * G + g] * (double)wscales[(size_t)n * G + g], bf16);
                        a = f64_to_ht((double)ph * (double)sc + (double)a, bf16);
                    }
                    acc = (double)a;
                } else {
                    /* e2m1 x e2m1 products are multiples of 1/4 up to 36: a 16-wide group sums exactly in fp32 */
                    double s = 0.0;
                    const float *pa = adec + (size_t)m * K;
                    for (int g = 0; g < G; g++) {
                        float sg = 0.f;
                        for (int i = 0; i < 16; i++) sg += pa[g * 16 + i] * wrow[g * 16 + i];
                        s += (double)sg * ((double)as[(size_t)m * G + g] * (double)wscales[(size_t)n * G + g]);
                    }
                    float a32 = (float)s;
                    if (alpha != 1.0f) a32 = a32 * alpha;                           /* :341-349 */
                    acc = (double)f32_to_ht(a32, bf16);                             /* packed_fp32_to_fp16 :351 */
                }
                /* EpilogueBias<USE_BIAS, USE_SCALE> */
                if (wcscales && bias) acc = (double)f64_to_ht(acc * (double)wcscales[n] + (double)bias[n], bf16);
                else if (wcscales) acc = (double)f64_to_ht(acc * (double)wcscales[n], bf16);
                else if (bias) acc = (double)f64_to_ht(acc + (double)bias[n], bf16);
                /* EpilogueLoraUp: hT(lora_act * scale) x lora_up accumulated, fp32 psum -> hT */
                if (R > 0 && lora_up) {
                    double add = 0.0;
                    for (int r = 0; r < R; r++)
                        add += (double)f32_to_ht(la[(size_t)m * R + r] * 1.0f, bf16) * (double)lora_up[(size_t)n * R + r];
                    acc = (double)f32_to_ht((float)(acc + add), bf16);
                }
                float o = (float)acc;
                if (!bf16) {                                                        /* gemm_base.cuh:688-696 */
                    if (o > 65504.f) o = 65504.f;
                    if (o < -65504.f) o = -65504.f;
                }
                out[(size_t)m * N + n] = o;
            }
        }
        free(wrow);
    }
    free(qa);
    free(as);
    free(la);
    free(adec);
}

/* OpenMP team size for the calls above (torchrun exports OMP_NUM_THREADS=1, which would serialise the CPU arm) */
#include <omp.h>
void svdq_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
