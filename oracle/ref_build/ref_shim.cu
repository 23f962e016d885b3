// TEST INFRASTRUCTURE (not product code).
//
// extern "C" entry points over the reference's OWN host functions and module class, so that Python (ctypes)
// and the C++ twin test can drive the UNMODIFIED reference code on raw device pointers:
//
//   nref_quantize_w4a4_act_fuse_lora -> nunchaku::kernels::quantize_w4a4_act_fuse_lora  (src/kernels/zgemm/zgemm.h:39-46)
//   nref_gemm_w4a4                   -> nunchaku::kernels::gemm_w4a4                    (zgemm.h:8-36)
//   nref_linearattn_vk_mul_q         -> nunchaku::kernels::linearattn_vk_mul_q          (zgemm.h:37)
//   nref_attention_fp16              -> nunchaku::kernels::attention_fp16               (zgemm.h:70-74)
//   nref_test_rmsnorm_rope / nref_test_pack_qkv -> kernels::test_*                      (zgemm.h:80-81)
//   nref_gemv_awq                    -> gemv_awq                                        (src/kernels/awq/gemv_awq.h)
//   nref_glue_*                      -> Silu/GELU::forward, LayerNorm/RMSNorm kernels, kernels::{add,mul_add_batch,cast,split_mod}
//   nref_linear_*                    -> class GEMM_W4A4                                  (src/Linear.h:53-120, Linear.cpp:90-502)
//
// This file is compiled twice by oracle/ref_build/build_ref.sh, with the reference's headers on the include path:
//   * into oracle/_ref/libnunchaku_ref.so together with the reference's own kernel objects (the GPU oracle and
//     the `reference_gpu` bench leg), and
//   * into oracle/_ref/libnunchaku_seam.so together with the reference's src/Linear.cpp, src/Module.cpp,
//     src/activation.cpp, src/layernorm.cpp objects and OUR forwarding definitions of the zgemm.h functions
//     (nunchaku_b200/csrc/seam/zgemm_b200.cpp) -- the proof that the reference's C++ host layer links and runs
//     unchanged on top of libnunchaku_b200.so (SURVEY section 8 rows a4 / b).
// Nothing here is copied from the reference; it only calls its public functions.
#include <cstring>
#include <string>

#include "Linear.h"
#include "Module.h"
#include "Tensor.h"
#include "activation.h"
#include "common.h"
#include "kernels/activation_kernels.h"
#include "kernels/awq/gemv_awq.h"
#include "kernels/layernorm_kernels.h"
#include "kernels/misc_kernels.h"
#include "kernels/zgemm/zgemm.h"
#include "layernorm.h"

extern "C" {

// dtype codes == the reference's Tensor::ScalarType enumerators (src/Tensor.h:215-226)
typedef struct nref_tensor {
    void *ptr;     // NULL == absent (the reference's default-constructed Tensor{})
    int dtype;     // 1 int8, 2 int16, 3 int32, 4 int64, 5 fp16, 6 fp32, 7 bf16, 8 fp8_e4m3, 9 fp8_e5m2
    int ndim;      // <= 5
    int shape[5];
    int on_cpu;    // 1: host memory (wtscale)
} nref_tensor;

const char *nref_last_error(void);
}

#ifdef NREF_SEAM_BUILD
namespace nunchaku::kernels {
void b200_invalidate_all();  // nunchaku_b200/csrc/seam/zgemm_b200.cpp: what a maintainer calls from GEMM_W4A4::loadParam
}
#endif

namespace {

thread_local std::string g_err;

// a Buffer over memory owned by the caller (torch)
class BufferExternal : public Buffer {
public:
    BufferExternal(void *p, size_t bytes, bool cpu) {
        this->ptr = p;
        this->size = bytes;
        if (cpu) {
            this->device = Device::cpu();
        } else {
            int d = 0;
            cudaGetDevice(&d);
            this->device = Device::cuda(d);
        }
    }
    bool isAsyncBuffer() override { return true; }  // never locked: the caller keeps it alive
};

Tensor wrap(const nref_tensor *t) {
    Tensor r;
    if (!t || !t->ptr || t->ndim <= 0) return r;
    std::vector<int> shape(t->shape, t->shape + t->ndim);
    r.shape = TensorShape(shape);
    r.scalarType = static_cast<Tensor::ScalarType>(t->dtype);
    r.buffer = std::make_shared<BufferExternal>(t->ptr, r.shape.size() * Tensor::scalarSize.at(r.scalarType), t->on_cpu != 0);
    return r;
}

// copy a tensor the reference allocated (cudaMallocAsync) into the caller's buffer
void copy_out(const nref_tensor *dst, Tensor src) {
    Tensor d = wrap(dst);
    if (!d.valid()) throw std::invalid_argument("output tensor missing");
    if (d.numel() * d.scalar_size() != src.numel() * src.scalar_size()) throw std::invalid_argument("output size mismatch: reference produced " + src.shape.str());
    checkCUDA(cudaMemcpyAsync(d.data_ptr(), src.data_ptr(), src.numel() * src.scalar_size(), cudaMemcpyDeviceToDevice, getCurrentCUDAStream()));
}

template <typename F>
int guarded(void *stream, F &&f) {
    try {
        CUDAStreamContext sctx(static_cast<cudaStream_t>(stream));
        f();
        checkCUDA(cudaGetLastError());
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

struct OneTensorProvider : TensorsProvider {
    std::string key;
    Tensor t;
    bool contains(const std::string &k) const override { return k == key; }
    Tensor getTensor(const std::string &k) override { return k == key ? t : Tensor{}; }
};

}  // namespace

extern "C" {

const char *nref_last_error(void) { return g_err.c_str(); }

int nref_quantize_w4a4_act_fuse_lora(const nref_tensor *input, const nref_tensor *output, const nref_tensor *oscales,
                                     const nref_tensor *lora_down, const nref_tensor *lora_act_out, const nref_tensor *smooth,
                                     int fuse_glu, int fp4, void *stream) {
    return guarded(stream, [&] {
        nunchaku::kernels::quantize_w4a4_act_fuse_lora(wrap(input), wrap(output), wrap(oscales), wrap(lora_down), wrap(lora_act_out),
                                                       wrap(smooth), fuse_glu != 0, fp4 != 0);
    });
}

// argument order == zgemm.h:8-36
int nref_gemm_w4a4(const nref_tensor *act, const nref_tensor *wgt, const nref_tensor *out, const nref_tensor *qout,
                   const nref_tensor *ascales, const nref_tensor *wscales, const nref_tensor *oscales, const nref_tensor *poolout,
                   const nref_tensor *lora_act_in, const nref_tensor *lora_up, const nref_tensor *lora_down,
                   const nref_tensor *lora_act_out, const nref_tensor *norm_q, const nref_tensor *norm_k,
                   const nref_tensor *rotary_emb, const nref_tensor *bias, const nref_tensor *smooth_factor,
                   const nref_tensor *out_vk, const nref_tensor *out_linearattn, int act_unsigned, const float *lora_scales,
                   int n_lora_scales, int fuse_silu, int fp4, float alpha, const nref_tensor *wcscales, const nref_tensor *out_q,
                   const nref_tensor *out_k, const nref_tensor *out_v, int attn_tokens, void *stream) {
    return guarded(stream, [&] {
        std::vector<float> ls(lora_scales, lora_scales + n_lora_scales);
        nunchaku::kernels::gemm_w4a4(wrap(act), wrap(wgt), wrap(out), wrap(qout), wrap(ascales), wrap(wscales), wrap(oscales),
                                     wrap(poolout), wrap(lora_act_in), wrap(lora_up), wrap(lora_down), wrap(lora_act_out), wrap(norm_q),
                                     wrap(norm_k), wrap(rotary_emb), wrap(bias), wrap(smooth_factor), wrap(out_vk),
                                     wrap(out_linearattn), act_unsigned != 0, ls, fuse_silu != 0, fp4 != 0, alpha, wrap(wcscales),
                                     wrap(out_q), wrap(out_k), wrap(out_v), attn_tokens);
    });
}

int nref_linearattn_vk_mul_q(const nref_tensor *q, const nref_tensor *vk, void *stream) {
    return guarded(stream, [&] { nunchaku::kernels::linearattn_vk_mul_q(wrap(q), wrap(vk)); });
}

// (both builds: in the seam library `attention_fp16` is OUR definition and consumes OUR PackQKV layout)
int nref_attention_fp16(const nref_tensor *q, const nref_tensor *k, const nref_tensor *v, const nref_tensor *o, float scale, void *stream) {
    return guarded(stream, [&] { nunchaku::kernels::attention_fp16(wrap(q), wrap(k), wrap(v), wrap(o), scale); });
}
#ifndef NREF_SEAM_BUILD

int nref_test_rmsnorm_rope(const nref_tensor *input, const nref_tensor *output, const nref_tensor *norm_q, const nref_tensor *norm_k,
                           const nref_tensor *rotary_emb, void *stream) {
    return guarded(stream, [&] { nunchaku::kernels::test_rmsnorm_rope(wrap(input), wrap(output), wrap(norm_q), wrap(norm_k), wrap(rotary_emb)); });
}

int nref_test_pack_qkv(const nref_tensor *input, const nref_tensor *out_q, const nref_tensor *out_k, const nref_tensor *out_v, int num_tokens,
                       void *stream) {
    return guarded(stream, [&] { nunchaku::kernels::test_pack_qkv(wrap(input), wrap(out_q), wrap(out_k), wrap(out_v), num_tokens); });
}

#endif

// (both builds: in the seam library `gemv_awq` is OUR definition, nunchaku_b200/csrc/seam/awq_b200.cpp)
int nref_gemv_awq(const nref_tensor *x, const nref_tensor *qweight, const nref_tensor *scales, const nref_tensor *zeros, int m, int n, int k,
                  int group_size, const nref_tensor *out, void *stream) {
    return guarded(stream, [&] { copy_out(out, gemv_awq(wrap(x), wrap(qweight), wrap(scales), wrap(zeros), m, n, k, group_size)); });
}

// ---- glue ----------------------------------------------------------------------------------
// kind: 0 silu, 1 gelu_new   (Silu::forward / GELU::forward, src/activation.cpp:4-14)
int nref_glue_activation(int kind, const nref_tensor *x, const nref_tensor *out, void *stream) {
    return guarded(stream, [&] { copy_out(out, kind == 0 ? Silu::forward(wrap(x)) : GELU::forward(wrap(x))); });
}
int nref_glue_layernorm(const nref_tensor *x, const nref_tensor *weight, const nref_tensor *bias, const nref_tensor *out, float eps, void *stream) {
    return guarded(stream, [&] { layernorm_general(wrap(out), wrap(x), wrap(weight), wrap(bias), eps); });
}
int nref_glue_rms_norm(const nref_tensor *x, const nref_tensor *weight, const nref_tensor *out, float eps, void *stream) {
    return guarded(stream, [&] {
        Tensor o = wrap(out), i = wrap(x), w = wrap(weight);
        rms_norm(o, i, w, eps, false);
    });
}
int nref_glue_add(const nref_tensor *a, const nref_tensor *b, const nref_tensor *out, void *stream) {
    return guarded(stream, [&] { copy_out(out, nunchaku::kernels::add(wrap(a), wrap(b))); });
}
int nref_glue_mul_add_batch(const nref_tensor *x, const nref_tensor *scale, int batch_scale, double scale_shift, const nref_tensor *bias,
                            int batch_bias, void *stream) {
    return guarded(stream, [&] { nunchaku::kernels::mul_add_batch(wrap(x), wrap(scale), batch_scale != 0, scale_shift, wrap(bias), batch_bias != 0); });
}
int nref_glue_cast(const nref_tensor *in, const nref_tensor *out, void *stream) {
    return guarded(stream, [&] { nunchaku::kernels::cast(wrap(in), wrap(out)); });
}
int nref_glue_split_mod(const nref_tensor *in, const nref_tensor *outs, int n, void *stream) {
    return guarded(stream, [&] {
        auto emit = [&](auto arr) {
            for (int i = 0; i < n; i++) copy_out(&outs[i], arr[i]);
        };
        switch (n) {
            case 2: emit(nunchaku::kernels::split_mod<2>(wrap(in))); break;
            case 3: emit(nunchaku::kernels::split_mod<3>(wrap(in))); break;
            case 4: emit(nunchaku::kernels::split_mod<4>(wrap(in))); break;
            case 5: emit(nunchaku::kernels::split_mod<5>(wrap(in))); break;
            case 6: emit(nunchaku::kernels::split_mod<6>(wrap(in))); break;
            default: throw std::invalid_argument("split_mod: n in 2..6");
        }
    });
}

// ---- class GEMM_W4A4 (src/Linear.h:53-120) ------------------------------------------------------------------------------
void *nref_linear_create(int in_features, int out_features, int bias, int fp4, int dtype) {
    try {
        int dev = 0;
        checkCUDA(cudaGetDevice(&dev));
        return new GEMM_W4A4(in_features, out_features, bias != 0, fp4 != 0, static_cast<Tensor::ScalarType>(dtype), Device::cuda(dev));
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void nref_linear_destroy(void *h) { delete static_cast<GEMM_W4A4 *>(h); }

// key in {qweight, wscales, bias, lora_down, lora_up, smooth, wtscale, wcscales}: goes through Module::loadParams ->
// GEMM_W4A4::loadParam (src/Linear.cpp:124-154), i.e. the reference's own per-key rules (re-allocation on rank change etc.)
int nref_linear_load(void *h, const char *key, const nref_tensor *src, void *stream) {
    return guarded(stream, [&] {
        OneTensorProvider p;
        p.key = key;
        p.t = wrap(src);
        static_cast<GEMM_W4A4 *>(h)->loadParams(p, /*partial=*/true);
#ifdef NREF_SEAM_BUILD
        nunchaku::kernels::b200_invalidate_all();  // parameters changed in place: drop the converted copies
#endif
    });
}
int nref_linear_lora_rank(void *h) { return static_cast<GEMM_W4A4 *>(h)->lora_rank; }
int nref_linear_set_lora_scales(void *h, const float *s, int n) {
    auto *m = static_cast<GEMM_W4A4 *>(h);
    m->lora_scales.assign(s, s + n);
    return 0;
}
// fuse: 0 = forward(x), 2 = forward_silu(x)     (FuseOptions, Linear.h:55-59)
int nref_linear_forward(void *h, const nref_tensor *x, const nref_tensor *out, int fuse, void *stream) {
    return guarded(stream, [&] {
        auto *m = static_cast<GEMM_W4A4 *>(h);
        copy_out(out, fuse == 2 ? m->forward_silu(wrap(x)) : m->forward(wrap(x)));
    });
}
// fc1.forward(x, GELU_QUANT, fc2) -> fc2.forward_quant(qact): the fused MLP of FluxModel.cpp:355-359 / :561-567
int nref_linear_forward_mlp(void *h_fc1, void *h_fc2, const nref_tensor *x, const nref_tensor *out, void *stream) {
    return guarded(stream, [&] {
        auto *fc1 = static_cast<GEMM_W4A4 *>(h_fc1);
        auto *fc2 = static_cast<GEMM_W4A4 *>(h_fc2);
        auto q = std::get<GEMM_W4A4::QuantizedActivation>(fc1->forward(wrap(x), GEMM_W4A4::FuseOptions::GELU_QUANT, fc2));
        copy_out(out, fc2->forward_quant(q));
    });
}
// the 10-argument forward (QKV projection with RMSNorm + RoPE [+ PackQKV]), Linear.cpp:169-268
int nref_linear_forward_qkv(void *h, const nref_tensor *x, const nref_tensor *out, const nref_tensor *norm_q, const nref_tensor *norm_k,
                            const nref_tensor *rotary_emb, const nref_tensor *out_q, const nref_tensor *out_k, const nref_tensor *out_v,
                            int num_tokens, void *stream) {
    return guarded(stream, [&] {
        static_cast<GEMM_W4A4 *>(h)->forward(wrap(x), wrap(out), {}, wrap(norm_q), wrap(norm_k), wrap(rotary_emb), wrap(out_q), wrap(out_k),
                                             wrap(out_v), num_tokens);
    });
}
// quantize only: act / ascales / lora_act copied out in the implementation's own inter-op layout
int nref_linear_quantize(void *h, const nref_tensor *x, const nref_tensor *act, const nref_tensor *ascales, const nref_tensor *lora_act,
                         int fuse_glu, void *stream) {
    return guarded(stream, [&] {
        auto q = static_cast<GEMM_W4A4 *>(h)->quantize(wrap(x), fuse_glu != 0);
        copy_out(act, q.act);
        copy_out(ascales, q.ascales);
        if (q.lora_act.numel() > 0) copy_out(lora_act, q.lora_act);
    });
}

}  // extern "C"

// ---- out-of-scope kernels src/Linear.cpp names (W8A8, CUTLASS fp16 GEMM, depth-wise conv): not built, never called here ------
#ifndef NREF_HAVE_OUT_OF_SCOPE
Tensor gemm_f16(Tensor, Tensor, Tensor, Tensor, float) { throw std::runtime_error("gemm_f16: out of scope, not built"); }
Tensor dwconv_f16(Tensor, Tensor, Tensor, Tensor) { throw std::runtime_error("dwconv_f16: out of scope, not built"); }
namespace nunchaku::kernels {
void gemm_w8a8(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor) { throw std::runtime_error("gemm_w8a8: out of scope, not built"); }
void quantize_w8a8_act(Tensor, Tensor, Tensor, bool) { throw std::runtime_error("quantize_w8a8_act: out of scope, not built"); }
}  // namespace nunchaku::kernels
#endif
