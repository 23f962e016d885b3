/*
 * nunchaku_b200 -- C ABI of the B200-native SVDQuant W4A4 fused linear.
 *
 * This is the drop-in boundary.  Every entry point takes plain device
 * pointers, sizes and a CUDA stream handle -- no torch / C++ types -- and returns 0 on
 * success or a negative nb200_status; nb200_last_error() returns a thread-local message.
 * All launches are asynchronous on `stream` (the reference launches on torch's current
 * stream: src/interop/torch.cpp:84-91, src/common.h:60-67).
 *
 * What each entry point replaces in the reference (/root/reference):
 *
 *   nb200_quantize_w4a4_act_fuse_lora  <->  nunchaku::kernels::quantize_w4a4_act_fuse_lora
 *        src/kernels/zgemm/zgemm.h:39-46, gemm_w4a4.cu:113-125,
 *        gemm_w4a4_launch_impl.cuh:451-521, kernel gemm_w4a4.cuh:1097-1184
 *   nb200_gemm_w4a4                    <->  nunchaku::kernels::gemm_w4a4
 *        src/kernels/zgemm/zgemm.h:8-36, gemm_w4a4.cu:34-105,
 *        gemm_w4a4_launch_impl.cuh:7-424, kernels gemm_w4a4.cuh:358-405,1046-1095
 *   nb200_repack_*                     <->  (no reference twin) one-time, at parameter-load
 *        time, conversion of the reference's mma.sync-fragment-ordered tensors
 *        (nunchaku/lora/flux/packer.py:187-437; GEMM_W4A4::loadParam src/Linear.cpp:124-154)
 *        into the TMA / tcgen05 friendly layouts below.  Pure permutations (+ dtype
 *        widening for bias/scales, + folding 1/(alpha*wcscale) into lora_up).
 *
 * Layouts ("B200 layouts", all row-major, K innermost):
 *
 *   act / qweight  INT4 : u8 [rows, K/2].  Each aligned group of 8 consecutive k is one
 *                         little-endian u32; nibble p (p<4) holds element 2p, nibble p+4
 *                         holds element 2p+1.  Signed values are stored offset-binary
 *                         (q + 8); unsigned activations (act_unsigned) are stored as is.
 *                  NVFP4: u8 [rows, K/2], e2m1 codes, low nibble = even k.
 *   ascales/wscales INT4 : hT [K/64, rows]            (group-major, rows contiguous)
 *                  NVFP4: u8 (ue4m3) tiles [rows/128][K/64][32][16]: the byte for
 *                         (row r, 16-group c of a 64-wide k block) sits at
 *                         (r%32)*16 + ((r%128)/32)*4 + c  -- the tcgen05.cp 32x128b layout.
 *   lora_act            : f32 [Mp, R] row-major.
 *   lora_up  (B200)     : hT, blocks [Rp/32][N/8][4][8][8]  (UMMA no-swizzle K-major core
 *                         matrices; Rp = R rounded up to 32, zero padded), pre-multiplied by
 *                         1/cscale[n].
 *   lora_down (B200)    : hT, [K/32][Rp/8][32 lanes][8]  mma.sync B-fragment order used by
 *                         the quantize kernel (see csrc/quantize.cu); for the fused fc1 epilogue
 *                         the next layer's factor is plain row-major [R, K].
 *   bias, cscale        : f32 [N]    (cscale = alpha * wcscales, 1 when absent)
 *   smooth              : hT [K] in natural order.
 */
#ifndef NUNCHAKU_B200_H_
#define NUNCHAKU_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB200_ABI_VERSION 2

typedef enum nb200_status {
    NB200_OK = 0,
    NB200_ERR_INVALID_ARGUMENT = -1, /* shape / alignment / null-pointer precondition  */
    NB200_ERR_UNSUPPORTED = -2,      /* valid in the reference, not built here yet      */
    NB200_ERR_CUDA = -3,             /* a CUDA runtime / driver call failed             */
    NB200_ERR_ARCH = -4              /* device is not sm_100                            */
} nb200_status;

typedef enum nb200_dtype { NB200_FP16 = 0, NB200_BF16 = 1, NB200_FP32 = 2 /* glue ops only */ } nb200_dtype;

/* mid-epilogue activation (reference: EpilogueNop / EpilogueSilu / EpilogueGelu) */
typedef enum nb200_act { NB200_ACT_NONE = 0, NB200_ACT_SILU = 1, NB200_ACT_GELU = 2 } nb200_act;

#define NB200_MAX_LORA_SCALES 64 /* MAX_RANK / 16, src/kernels/zgemm/lora.cuh:22,41 */

/* ---- introspection ------------------------------------------------------------------ */
int nb200_abi_version(void);
const char *nb200_last_error(void);
/* 0 when the current device is compute capability 10.x, NB200_ERR_ARCH otherwise. */
int nb200_check_device(void);

/* ---- one-time repack of reference-layout parameters (device -> device) --------------- */
/* qweight: reference int8 [N, K/2] (packer.py:187-239)  ->  B200 u8 [N, K/2]. */
int nb200_repack_qweight(const void *src, void *dst, int N, int K, int fp4, void *stream);
/* INT4 wscales: reference hT "[K/64, N]" (packer.py:241-301) -> hT [K/64, N] natural. */
int nb200_repack_wscales_int4(const void *src, void *dst, int N, int K, void *stream);
/* NVFP4 wscales: reference fp8 "[K/16, N]" (packer.py:303-360) -> ue4m3 tiles. */
int nb200_repack_wscales_fp4(const void *src, void *dst, int N, int K, void *stream);
/* bias / smooth / wcscales: reference hT [N] in pack_scale(group_size=-1) order ->
 * natural order; out_f32 != 0 widens to float and multiplies by `mul` (cscale = alpha*wcs). */
int nb200_repack_channel_vector(const void *src, void *dst, int N, int dtype, int out_f32, float mul, void *stream);
/* lora_up: reference hT [N, R] (packer.py:362-398) -> B200 blocks, divided by cscale[n]
 * (cscale may be NULL == 1).  dst holds N * Rp elements, Rp = ceil(R/32)*32. */
int nb200_repack_lora_up(const void *src, void *dst, const float *cscale, int N, int R, int dtype, void *stream);
/* lora_down: reference hT [K, R] -> B200 fragment order for the quantize kernel
 * (dst holds 2 * K * Rp elements, Rp = ceil(R/32)*32: the TMA kernel's true-k-order fragments
 * followed by the register-streaming (GLU) kernel's k-permuted fragments). */
int nb200_repack_lora_down(const void *src, void *dst, int K, int R, int dtype, void *stream);
/* lora_down of the NEXT layer for the fused fc1 epilogue: reference hT [K, R] -> logical [R, K]
 * row-major (TMA source; K here is the fc1 output width N). */
int nb200_repack_lora_down_next(const void *src, void *dst, int K, int R, int dtype, void *stream);

/* ---- activation quantize + low-rank down projection ---------------------------------- */
typedef struct nb200_quantize_args {
    const void *input;     /* hT [M, K] row-major ([M, 2K] when fuse_glu)                 */
    void *output;          /* u8 [Mp, K/2]                                                 */
    void *oscales;         /* INT4: hT [K/64, Mp];  NVFP4: ue4m3 tiles, K/16*Mp bytes      */
    const void *lora_down; /* B200 layout (nb200_repack_lora_down), R > 0                  */
    float *lora_act_out;   /* f32 [Mp, R]; fully overwritten (reference zero-fills first)  */
    const void *smooth;    /* hT [K] natural order, or NULL                                */
    int M;                 /* valid rows                                                   */
    int Mp;                /* padded rows, multiple of 256 (pad_size, ops/quantize.py:66)  */
    int K;                 /* multiple of 128                                              */
    int R;                 /* multiple of 16                                               */
    int dtype;             /* nb200_dtype                                                  */
    int fuse_glu;
    int fp4;
    /* optional scratch for the small-M split-K path (deterministic last-block reduction); NULL
     * disables the split.  nb200_quantize_workspace_bytes() gives a sufficient size.  Must be
     * zero-initialised once; the kernel leaves it clean. */
    void *workspace;
    long long workspace_bytes;
    /* INT4 only, not in the reference's quantize op: quantise (x + 0.171875) / smooth to UNSIGNED codes (scale = max / 15),
     * i.e. exactly what the reference's fused GELU epilogue produces for the next layer (EpilogueQuantize<false, true>,
     * gemm_w4a4_launch_impl.cuh:282-310); the low-rank projection still sees the unshifted input.  Lets a caller split
     * "GEMM + GELU" and "quantise for fc2" into two launches without changing the numbers.                            */
    int act_unsigned_shift;
} nb200_quantize_args;

long long nb200_quantize_workspace_bytes(int Mp, int K);

int nb200_quantize_w4a4_act_fuse_lora(const nb200_quantize_args *args, void *stream);

/* ---- fused W4A4 GEMM ------------------------------------------------------------------ */
typedef struct nb200_gemm_args {
    /* operands (B200 layouts) */
    const void *act;       /* u8 [Mp, K/2]                                                 */
    const void *wgt;       /* u8 [N, K/2]                                                  */
    const void *ascales;
    const void *wscales;
    /* epilogue inputs */
    const float *bias;     /* f32 [N] or NULL                                              */
    const float *cscale;   /* f32 [N] or NULL (== alpha * wcscales)                        */
    const float *lora_act_in; /* f32 [Mp, R_up] or NULL                                    */
    const void *lora_up;   /* B200 layout or NULL                                          */
    /* outputs: mode is selected by which are non-NULL, as in the reference
     * (gemm_w4a4_launch_impl.cuh:282,311,347,407)                                         */
    void *out;             /* hT [M_out, N_out] row-major or NULL                          */
    void *qout;            /* u8 [Mp, N/2]: next layer's quantized activations, or NULL    */
    void *oscales;         /* scales of qout                                               */
    const void *smooth_next;    /* hT [N] natural, required with qout                      */
    const void *lora_down_next; /* hT [R_down, N] row-major (nb200_repack_lora_down_next)      */
    float *lora_act_out;   /* f32 [Mp, R_down]; zeroed then accumulated                    */
    const void *norm_q;    /* hT [128]   RMSNorm weights (rotary mode)                     */
    const void *norm_k;
    const float *rotary_emb; /* f32, reference pack_rotemb layout [Mp, 128]                */
    /* sizes */
    int Mp, N, K;          /* Mp % 256 == 0, N % 128 == 0, K % 128 == 0                    */
    int M_out, N_out;      /* rows / cols of `out` actually stored (<= Mp, N)              */
    int R_up, R_down;
    /* flags */
    int dtype;             /* nb200_dtype of out / scales / lora                           */
    int fp4;
    int act_unsigned;      /* INT4 only: act nibbles are 0..15                             */
    int mid_act;           /* nb200_act; GELU is implied when qout != NULL                 */
    float lora_scales[NB200_MAX_LORA_SCALES];
    /* tuning (0 = auto) */
    int block_n;           /* 128 / 256: single-CTA tiles; 512: CTA-pair (cta_group::2) 256x256 tiles      */
    int num_sms;
    void *prof;            /* optional device buffer, 16 x int64 per CTA: barrier-wait cycle counters  */
    /* rotary mode, optional: attention-ready Q / K / V instead of `out` (EpiloguePackQKV,
     * src/kernels/zgemm/epilogues.cuh:427-550; wiring gemm_w4a4_launch_impl.cuh:376-393).
     * fp16 [heads, rows >= Mp, 128] each, row pitch 128, head pitch stride_head_* elements; element order is
     * plain row-major (the reference writes its own attention kernel's fragment order).  Rows >= attn_tokens are
     * filled with 0 (Q, V) / NaN (K) exactly like the reference's mask (epilogues.cuh:479-489, 539-545).
     * `out` may be given as well: then it is a SCRATCH of exactly [Mp, N] hT (M_out = Mp, N_out = N) whose contents are
     * unspecified afterwards; the NVFP4 cluster route uses it to run the plain GEMM followed by the RMSNorm + RoPE + pack
     * kernel instead of the fused epilogue (same bits, ~30 % faster at FLUX sizes).                              */
    void *out_q, *out_k, *out_v;
    long long stride_head_q, stride_head_k, stride_head_v;
    int attn_tokens;
    /* fused quantise epilogue with R_down > 0, optional: device scratch of >= nb200_gemm_workspace_bytes(Mp, R_down) bytes.  With it the per-CTA
     * partial projections of lora_act_out go to the scratch and a small second kernel adds them in a fixed order -- identical bits on every launch,
     * no zero-fill of lora_act_out; without it (NULL / too small) they are added with fp32 atomics like the reference's red.global.add.f32
     * (lora.cuh:320-353), which is correct to fp32 rounding but not run-to-run bit-stable.  One workspace per concurrently running launch. */
    void *workspace;
    long long workspace_bytes;
    /* SANA linear attention as a GEMM epilogue (EpilogueLiteLA, src/kernels/zgemm/epilogues.cuh:552-691; wiring gemm_w4a4_launch_impl.cuh:311-346),
     * optional: out_vk f32 [Mp / vk_tokens, N / 96, 33, 32] (zero-filled inside the call, accumulated with fp32 atomics like the reference's
     * reduce_add) and `out` = relu(Q) hT [Mp, N / 3] (M_out = Mp, N_out = N / 3).  Channel layout of the projection: [ Q (N/3) | per head: K (32),
     * V (32) ]; vk[v][k] = sum_t V[t,v] relu(K[t,k]), vk[32][k] = sum_t relu(K[t,k]) over the vk_tokens rows of an image.  Needs (N / 3) % 128 == 0
     * (the reference asserts numBlocksN % 3 == 0) and vk_tokens % 128 == 0; other shapes: plain GEMM + nb200_litela_vk below.                      */
    float *out_vk;
    int vk_tokens;
} nb200_gemm_args;

int nb200_gemm_w4a4(const nb200_gemm_args *args, void *stream);
long long nb200_gemm_workspace_bytes(int Mp, int R_down);

/* ---- elementwise / row-reduce glue between the linears ------------------
 * All tensors contiguous, 16-byte aligned, on the current device; dtype is an nb200_dtype.            */

/* out = act(x);  kind = NB200_ACT_SILU (replaces Silu::forward, src/activation.cpp:4-8 -> vllm silu,
 * kernels/activation_kernels_impl.cuh:7-10) or NB200_ACT_GELU (GELU::forward -> gelu_new, :93-97).        */
int nb200_activation(int kind, int dtype, const void *x, void *out, long long numel, void *stream);

/* out = LayerNorm(x) over the last dim, optional affine (weight / bias may be NULL); fp16 / bf16.
 * Replaces LayerNorm::forward -> layernorm_general (src/layernorm.cpp:14-18, kernels/layernorm_kernels.cu:36-58). */
int nb200_layernorm(int dtype, const void *x, const void *weight, const void *bias, void *out, long long rows, int hidden,
                    float eps, void *stream);
/* LayerNorm + AdaLN modulation in one pass: out = LN(x) * (mod_scale[c] + scale_shift) + mod_shift[c] with mod_scale / mod_shift hT [hidden]
 * (one modulation vector for all rows); same rounding points as nb200_layernorm followed by nb200_mul_add_batch (src/FluxModel.cpp:41-96,
 * src/kernels/misc_kernels.cu:70-131), i.e. bit-identical, with one read and one write of the activations instead of two. */
int nb200_layernorm_mod(int dtype, const void *x, const void *weight, const void *bias, const void *mod_scale, const void *mod_shift, float scale_shift,
                        void *out, long long rows, int hidden, float eps, void *stream);

/* out = T(x * rsqrt(mean(x^2) + eps)) * weight.  Replaces RMSNorm::forward (use_quant = false) -> rms_norm
 * (src/layernorm.cpp:20-24, kernels/layernorm_kernels.cu:6-34).                                          */
int nb200_rms_norm(int dtype, const void *x, const void *weight, void *out, long long rows, int hidden, float eps, void *stream);

/* out = a + b.  Replaces kernels::add (src/kernels/misc_kernels.cu:7-27).                                */
int nb200_add(int dtype, const void *a, const void *b, void *out, long long numel, void *stream);

/* In place, per batch b: x[b,i] = x[b,i] * (scale[b?, i % numel_scale] + scale_shift) + bias[b?, i % numel_bias]
 * (scale == NULL: x += bias).  `numel` is per batch; a stride of 0 shares scale / bias across the batch.
 * Replaces kernels::mul_add and kernels::mul_add_batch (src/kernels/misc_kernels.cu:29-131).               */
int nb200_mul_add_batch(int dtype, void *x, const void *scale, const void *bias, float scale_shift, int batch, long long numel,
                        long long numel_scale, long long numel_bias, long long stride_x, long long stride_scale,
                        long long stride_bias, void *stream);

/* De-interleave the last dim: outs[k][i] = input[i * n + k], n in 2..6.  Replaces kernels::split_mod<N>
 * (src/kernels/misc_kernels.cu:187-214).                                                                */
int nb200_split_mod(int dtype, const void *input, void *const *outs, int n, long long numel, void *stream);

/* Element type conversion between fp16 / bf16 / fp32 (fp16 results clamp to +-65504).  Replaces kernels::cast
 * (src/kernels/misc_kernels.cu:256-285).                                                                 */
int nb200_cast(int dtype_in, const void *input, int dtype_out, void *output, long long numel, void *stream);

/* ---- SANA linear attention ---------------------------------------------------
 * Second half of the reference's EpilogueLiteLA (src/kernels/zgemm/epilogues.cuh:552-691, wiring
 * gemm_w4a4_launch_impl.cuh:311-346), applied to the QKV projection's hT output qkv [batch, tokens, N],
 * N = 3 * heads * 32 laid out [Q | per head: K(32) V(32)]:
 *   out_q  [batch, tokens, N/3] hT   = relu(Q)
 *   out_vk [batch, heads, 33, 32] f32: vk[v][k] = sum_t V[t,v] * relu(K[t,k]),  vk[32][k] = sum_t relu(K[t,k])
 * out_vk is zero-filled inside the call (launch_impl:336) and accumulated with fp32 atomics.                */
int nb200_litela_vk(int dtype, const void *qkv, void *out_q, float *out_vk, int batch, int tokens, int N, void *stream);

/* In place: q[b,t,h,:] <- (q . vk[b,h,0..31,:]) / (q . vk[b,h,32,:] + eps), q hT [batch, tokens, heads, 32].
 * Replaces kernels::linearattn_vk_mul_q (gemm_w4a4_launch_impl.cuh:427-448, epilogues.cuh:693-760; eps = 1e-6). */
int nb200_linearattn_vk_mul_q(int dtype, void *q, const float *vk, int batch, int tokens, int heads, float eps, void *stream);

/* ---- AWQ W4A16 GEMV for the AdaLN modulation linears -----------------------------------------
 * out[m, n] = sum_k x[m, k] * (code[n, k] * scales[k/64, n] + zeros[k/64, n]),  1 <= M <= 8, group_size == 64.
 * Replaces gemv_awq (src/kernels/awq/gemv_awq.cu:241-294, GEMV_AWQ::forward src/Linear.cpp:56-86, nunchaku/ops/gemv.py:10-58).
 * qweight int32 [OC/4, IC/8*4] in the CHECKPOINT layout (read in place, no repack), scales / zeros hT [IC/64, OC], x hT [M, IC],
 * out hT [M, OC]; the reference's arithmetic (hT dequant FMA, hT product, fp32 accumulation). */
int nb200_gemv_awq(int dtype, const void *x, const void *qweight, const void *scales, const void *zeros, void *out, int M, int OC, int IC,
                   int group_size, void *stream);
/* The same with the two neighbours of the modulation GEMV folded in: x <- silu(x) before the product when
 * fuse_silu != 0 (kernels::silu, src/kernels/activation_kernels_impl.cuh:7-10) and out += bias (hT [OC], may be NULL) after it -- bit-identical
 * to the three launches of AdaLayerNormZero::forward (src/FluxModel.cpp:41-96). */
int nb200_gemv_awq_fused(int dtype, const void *x, const void *qweight, const void *scales, const void *zeros, const void *bias, void *out, int M, int OC,
                         int IC, int group_size, int fuse_silu, void *stream);

/* ---- SANA GLUMBConv: depthwise 3x3 convolution, NHWC, stride 1, zero padding 1 ---------------------------
 * Replaces dwconv_f16 (src/kernels/dwconv.h:9, src/kernels/dwconv.cu:202-340; module DWCONV, src/Linear.cpp:541-551).
 * x / out hT [N, H, W, C] contiguous, weight hT [C, 3, 3] (= the reference's [C, 3, 3, 1]), bias hT [C] or NULL; C % 8 == 0. */
int nb200_dwconv3x3(int dtype, const void *x, const void *weight, const void *bias, void *out, int N, int H, int W, int C, void *stream);

/* ---- scaled-dot-product attention of the FLUX blocks ------------------------------------------------
 * o[b, i, h*128 + :] = softmax_j(scale * q[b,h,i,:] . k[b,h,j,:]) @ v[b,h,j,:], non-causal, head_dim 128.
 * Replaces nunchaku::kernels::attention_fp16 (src/kernels/zgemm/attention.cu:10-94, zgemm.h:70-74; nunchaku/csrc/ops.h attention_fp16).
 * q / k / v: fp16 [batch, heads, tokens, 128] contiguous, ROW-MAJOR inside a head -- what nb200_gemm_w4a4's PackQKV epilogue writes (the
 * reference keeps its mma.sync fragment order there).  Pad rows of K hold NaN and act as the key mask (NaN score -> -inf), pad rows of Q
 * are 0, of V 0, exactly as that epilogue produces them.  o: fp16 / bf16 (out_dtype) [batch, tokens_q, heads * 128].
 * tokens_q and tokens_kv must be multiples of 128 (the reference: 128 / 32).  Both GEMMs run on tcgen05 with fp32 accumulation. */
int nb200_attention_fp16(const void *q, const void *k, const void *v, void *o, int out_dtype, int batch, int heads, int tokens_q, int tokens_kv,
                         float scale, void *stream);

/* Number of kernels the last nb200_* call on this thread launched (bench bookkeeping). */
int nb200_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NUNCHAKU_B200_H_ */
