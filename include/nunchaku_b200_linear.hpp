// C++ twin of the SVDQuant W4A4 linear for host code written against the reference's C++ module
// surface:
//
//   reference                                   here
//   class GEMM_W4A4 : Module   src/Linear.h:53-120      nunchaku_b200::GEMM_W4A4
//   ctor / padded params       src/Linear.cpp:90-122    GEMM_W4A4::GEMM_W4A4
//   loadParam                  src/Linear.cpp:124-154   GEMM_W4A4::load_param
//   quantize                   src/Linear.cpp:426-462   GEMM_W4A4::quantize
//   forward / forward_silu /   src/Linear.cpp:156-424   GEMM_W4A4::forward / forward_silu /
//   forward(x, fuse, next) /                            forward_gelu_quant / forward_qkv / forward_quant
//   forward(x, out, pool, norm_q, ...) / forward_quant
//
// Header-only, C++17, depends only on the CUDA runtime and the C ABI (nunchaku_b200.h); every
// kernel launch goes through libnunchaku_b200.so.  Parameters are held exactly as the reference
// holds them (same names, shapes, dtypes, reference checkpoint layout -- FluxModel.cpp-style
// callers read / write them directly); the B200-layout copies the kernels consume are rebuilt
// lazily after a load_param().  Buffers are plain device allocations (the reference's ref-counted
// Tensor runtime is out of scope), tensors are passed as raw device pointers plus sizes.
// Errors throw std::runtime_error carrying nb200_last_error() (the reference: C++ exceptions for
// CUDA errors, assert-abort for preconditions).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "nunchaku_b200.h"

namespace nunchaku_b200 {

inline void cuda_check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
inline void nb_check(int status, const char *what) {
    if (status != NB200_OK) {
        const char *m = nb200_last_error();
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(status) + "): " + (m ? m : "?"));
    }
}

// Owning device allocation with a shape (row-major, contiguous).
class DeviceTensor {
public:
    DeviceTensor() = default;
    DeviceTensor(std::vector<int64_t> shape, size_t elem_size, bool zero = false) : shape_(std::move(shape)), elem_(elem_size) {
        size_t n = elem_;
        for (int64_t d : shape_) n *= static_cast<size_t>(d);
        bytes_ = n;
        if (bytes_ > 0) {
            void *p = nullptr;
            cuda_check(cudaMalloc(&p, bytes_), "cudaMalloc");
            ptr_ = std::shared_ptr<void>(p, [](void *q) { cudaFree(q); });
            if (zero) cuda_check(cudaMemset(p, 0, bytes_), "cudaMemset");
        }
    }
    void *data() const { return ptr_.get(); }
    template <typename T>
    T *data_ptr() const { return static_cast<T *>(ptr_.get()); }
    bool valid() const { return ptr_ != nullptr; }
    size_t bytes() const { return bytes_; }
    size_t numel() const { return elem_ ? bytes_ / elem_ : 0; }
    size_t elem_size() const { return elem_; }
    const std::vector<int64_t> &shape() const { return shape_; }

private:
    std::shared_ptr<void> ptr_;
    std::vector<int64_t> shape_;
    size_t elem_ = 0, bytes_ = 0;
};

class GEMM_W4A4 {
public:
    enum class FuseOptions { EMPTY = 0, GELU_QUANT, SILU };

    struct QuantizedActivation {
        DeviceTensor act;       // u8  [Mp, K/2]              (B200 inter-op layout)
        DeviceTensor ascales;   // hT [K/64, Mp] | ue4m3 scale tiles, K/16 * Mp bytes
        DeviceTensor lora_act;  // f32 [Mp, lora_rank]
        bool is_unsigned = false;
        int M = 0;              // valid rows
        int Mp = 0;             // rows padded to 256 (ops/quantize.py:66, launch_impl:462)
    };

    GEMM_W4A4(int in_features, int out_features, bool bias, bool use_fp4, nb200_dtype dtype, int device = 0)
        : in_features(in_features), out_features(out_features), in_features_pad((in_features + 127) / 128 * 128),
          out_features_pad((out_features + 127) / 128 * 128), use_fp4(use_fp4), lora_rank(0), dtype(dtype), device(device) {
        if (dtype != NB200_FP16 && dtype != NB200_BF16) throw std::invalid_argument("GEMM_W4A4: dtype must be fp16 or bf16");
        cuda_check(cudaSetDevice(device), "cudaSetDevice");
        nb_check(nb200_check_device(), "nb200_check_device");
        qweight = DeviceTensor({out_features_pad, in_features_pad / 2}, 1, true);
        wscales = use_fp4 ? DeviceTensor({in_features_pad / 16, out_features_pad}, 1, true)
                          : DeviceTensor({in_features_pad / 64, out_features_pad}, 2, true);
        if (bias) this->bias = DeviceTensor({out_features_pad}, 2, true);
        lora_down = DeviceTensor({in_features_pad, 0}, 2);
        lora_up = DeviceTensor({out_features_pad, 0}, 2);
        smooth = DeviceTensor({in_features_pad}, 2, true);
        wtscale = 1.0f;
    }

    // key in {qweight, wscales, bias, lora_down, lora_up, smooth, wtscale, wcscales}; `src` is a host or device pointer
    // (cudaMemcpyDefault) holding the tensor in the reference checkpoint layout.  lora_up / lora_down / wcscales are
    // re-allocated to the incoming shape like the reference does (Linear.cpp:124-141).
    void load_param(const std::string &key, const void *src, const std::vector<int64_t> &shape, size_t elem_size) {
        size_t bytes = elem_size;
        for (int64_t d : shape) bytes *= static_cast<size_t>(d);
        auto copy_into = [&](DeviceTensor &dst) {
            if (dst.bytes() != bytes) throw std::invalid_argument("load_param(" + key + "): size mismatch");
            if (bytes) cuda_check(cudaMemcpy(dst.data(), src, bytes, cudaMemcpyDefault), "cudaMemcpy");
        };
        if (key == "lora_down" || key == "lora_up") {
            if (shape.size() != 2) throw std::invalid_argument("load_param(" + key + "): 2-D tensor expected");
            DeviceTensor &dst = key == "lora_down" ? lora_down : lora_up;
            if (dst.shape() != shape) dst = DeviceTensor(shape, 2);
            copy_into(dst);
            lora_rank = static_cast<int>(shape[1]);
            lora_scales.resize((lora_rank + 15) / 16, 1.0f);
        } else if (key == "wcscales") {
            if (shape.size() != 1 || shape[0] != out_features_pad) throw std::invalid_argument("load_param(wcscales): [out_features_pad] expected");
            wcscales = DeviceTensor(shape, 2);
            copy_into(wcscales);
        } else if (key == "wtscale") {
            if (bytes != 4) throw std::invalid_argument("load_param(wtscale): one float32 expected");
            cuda_check(cudaMemcpy(&wtscale, src, 4, cudaMemcpyDefault), "cudaMemcpy");
        } else if (key == "qweight") {
            copy_into(qweight);
        } else if (key == "wscales") {
            copy_into(wscales);
        } else if (key == "bias") {
            if (!bias.valid()) throw std::invalid_argument("load_param(bias): the layer was built without bias");
            copy_into(bias);
        } else if (key == "smooth") {
            copy_into(smooth);
        } else {
            throw std::invalid_argument("load_param: unknown key " + key);
        }
        repacked_ = false;
    }

    // ---- reference: GEMM_W4A4::quantize (Linear.cpp:426-462) -------------------------------------------
    // x: hT [M, in_features]  ([M, 2 * in_features] with fuse_glu)
    QuantizedActivation quantize(const void *x, int M, bool fuse_glu, cudaStream_t stream = nullptr) {
        require_aligned_features();
        ensure_repacked(stream);
        QuantizedActivation q;
        q.M = M;
        q.Mp = (M + 255) / 256 * 256;
        q.act = DeviceTensor({q.Mp, in_features_pad / 2}, 1);
        q.ascales = use_fp4 ? DeviceTensor({in_features_pad / 16, q.Mp}, 1) : DeviceTensor({in_features_pad / 64, q.Mp}, 2);
        q.lora_act = DeviceTensor({q.Mp, lora_rank}, 4);
        q.is_unsigned = false;
        nb200_quantize_args a{};
        a.input = x;
        a.output = q.act.data();
        a.oscales = q.ascales.data();
        a.lora_down = b_lora_down_.data();
        a.lora_act_out = q.lora_act.data_ptr<float>();
        a.smooth = b_smooth_.data();
        a.M = M;
        a.Mp = q.Mp;
        a.K = in_features_pad;
        a.R = lora_rank;
        a.dtype = dtype;
        a.fuse_glu = fuse_glu;
        a.fp4 = use_fp4;
        const long long need = nb200_quantize_workspace_bytes(q.Mp, in_features_pad);
        if (need > 0 && static_cast<long long>(workspace_.bytes()) < need) workspace_ = DeviceTensor({need}, 1, true);
        a.workspace = workspace_.data();
        a.workspace_bytes = static_cast<long long>(workspace_.bytes());
        nb_check(nb200_quantize_w4a4_act_fuse_lora(&a, stream), "nb200_quantize_w4a4_act_fuse_lora");
        return q;
    }

    // ---- reference: Tensor forward(Tensor x) / forward_silu (Linear.cpp:156-162) ---------------------------
    // out: hT [M, out_features]
    void forward(const void *x, int M, void *out, cudaStream_t stream = nullptr) {
        forward_quant(quantize(x, M, false, stream), FuseOptions::EMPTY, nullptr, out, nullptr, stream);
    }
    void forward_silu(const void *x, int M, void *out, cudaStream_t stream = nullptr) {
        forward_quant(quantize(x, M, false, stream), FuseOptions::SILU, nullptr, out, nullptr, stream);
    }
    // ---- reference: forward(x, FuseOptions::GELU_QUANT, nextGEMM) (Linear.cpp:164-167, 272-331) --------------
    QuantizedActivation forward_gelu_quant(const void *x, int M, GEMM_W4A4 *next, cudaStream_t stream = nullptr) {
        QuantizedActivation qout;
        forward_quant(quantize(x, M, false, stream), FuseOptions::GELU_QUANT, next, nullptr, &qout, stream);
        return qout;
    }
    // ---- reference: forward_quant(qact) / forward_quant(qact, fuse, next) ------------------------------------
    void forward_quant(const QuantizedActivation &qact, FuseOptions fuse, GEMM_W4A4 *next, void *out, QuantizedActivation *qout,
                       cudaStream_t stream = nullptr) {
        ensure_repacked(stream);
        nb200_gemm_args a{};
        fill_common(a, qact);
        if (fuse == FuseOptions::GELU_QUANT) {
            if (next == nullptr || qout == nullptr) throw std::invalid_argument("GELU_QUANT needs the next GEMM and a qout");
            if (next->in_features_pad != out_features_pad) throw std::invalid_argument("next layer's in_features must match");
            next->ensure_repacked(stream);
            qout->M = qact.M;
            qout->Mp = qact.Mp;
            qout->act = DeviceTensor({qact.Mp, out_features_pad / 2}, 1);
            qout->ascales = use_fp4 ? DeviceTensor({out_features_pad / 16, qact.Mp}, 1) : DeviceTensor({out_features_pad / 64, qact.Mp}, 2);
            qout->lora_act = DeviceTensor({qact.Mp, next->lora_rank}, 4);
            qout->is_unsigned = !use_fp4;  // Linear.cpp:293: the shifted GELU output is non-negative
            a.qout = qout->act.data();
            a.oscales = qout->ascales.data();
            a.smooth_next = next->b_smooth_.data();
            a.R_down = next->lora_rank;
            if (next->lora_rank > 0) {
                a.lora_down_next = next->b_lora_down_next_.data();
                a.lora_act_out = qout->lora_act.data_ptr<float>();
            }
        } else {
            if (out == nullptr) throw std::invalid_argument("forward_quant: out is required");
            a.out = out;
            a.M_out = qact.M;
            a.N_out = out_features;
            a.mid_act = fuse == FuseOptions::SILU ? NB200_ACT_SILU : NB200_ACT_NONE;
        }
        nb_check(nb200_gemm_w4a4(&a, stream), "nb200_gemm_w4a4");
    }
    // ---- reference: forward(x, out, pool, norm_q, norm_k, rotary_emb, out_q, out_k, out_v, numTokens) -------------
    // norm_q / norm_k: hT [128]; rotary_emb: f32 in the reference's pack_rotemb layout [Mp, 128].  Either `out`
    // (hT [M, out_features]) or the three fp16 [heads, rows >= Mp, 128] tensors with their head pitches.
    void forward_qkv(const void *x, int M, void *out, const void *norm_q, const void *norm_k, const float *rotary_emb, void *out_q = nullptr,
                     void *out_k = nullptr, void *out_v = nullptr, long long stride_head_q = 0, long long stride_head_k = 0,
                     long long stride_head_v = 0, int num_tokens = 0, cudaStream_t stream = nullptr) {
        QuantizedActivation qact = quantize(x, M, false, stream);
        nb200_gemm_args a{};
        fill_common(a, qact);
        a.norm_q = norm_q;
        a.norm_k = norm_k;
        a.rotary_emb = rotary_emb;
        if (out_q != nullptr) {
            a.out_q = out_q;
            a.out_k = out_k;
            a.out_v = out_v;
            a.stride_head_q = stride_head_q;
            a.stride_head_k = stride_head_k;
            a.stride_head_v = stride_head_v;
            a.attn_tokens = num_tokens;
        } else {
            a.out = out;
            a.M_out = M;
            a.N_out = out_features;
        }
        nb_check(nb200_gemm_w4a4(&a, stream), "nb200_gemm_w4a4");
    }

public:
    const int in_features;
    const int out_features;
    const int in_features_pad;
    const int out_features_pad;
    const bool use_fp4;

    int lora_rank;
    std::vector<float> lora_scales;  // every 16 ranks share a scale (Linear.h:97)

    const nb200_dtype dtype;
    const int device;

    // reference-layout parameters, same names as Linear.h:106-117
    DeviceTensor qweight;    // i8  [out_pad, in_pad / 2]
    DeviceTensor wscales;    // hT [in_pad / 64, out_pad] | fp8 [in_pad / 16, out_pad]
    DeviceTensor bias;       // hT [out_pad] or invalid
    DeviceTensor lora_down;  // hT [in_pad, rank]
    DeviceTensor lora_up;    // hT [out_pad, rank]
    DeviceTensor smooth;     // hT [in_pad]
    float wtscale;           // host scalar (the reference keeps a CPU tensor, Linear.cpp:110-111)
    DeviceTensor wcscales;   // hT [out_pad] or empty

private:
    void require_aligned_features() const {
        if (in_features != in_features_pad) throw std::invalid_argument("GEMM_W4A4: in_features must be a multiple of 128");
        if (lora_rank <= 0 || lora_rank % 16 != 0) throw std::invalid_argument("GEMM_W4A4: load lora_down / lora_up (rank % 16 == 0) first");
    }
    void fill_common(nb200_gemm_args &a, const QuantizedActivation &qact) const {
        a.act = qact.act.data();
        a.wgt = b_qweight_.data();
        a.ascales = qact.ascales.data();
        a.wscales = b_wscales_.data();
        a.bias = b_bias_.valid() ? b_bias_.data_ptr<float>() : nullptr;
        a.cscale = b_cscale_.valid() ? b_cscale_.data_ptr<float>() : nullptr;
        a.lora_act_in = qact.lora_act.data_ptr<float>();
        a.lora_up = b_lora_up_.data();
        a.Mp = qact.Mp;
        a.N = out_features_pad;
        a.K = in_features_pad;
        a.R_up = lora_rank;
        a.dtype = dtype;
        a.fp4 = use_fp4;
        a.act_unsigned = qact.is_unsigned;
        for (int i = 0; i < NB200_MAX_LORA_SCALES; i++) a.lora_scales[i] = i < static_cast<int>(lora_scales.size()) ? lora_scales[i] : 0.f;
    }
    // B200-layout copies of the parameters (one repack per load; SURVEY.md Appendix A -> DESIGN.md section 2)
    void ensure_repacked(cudaStream_t stream) {
        if (repacked_) return;
        require_aligned_features();
        const int N = out_features_pad, K = in_features_pad, R = lora_rank, Rp = (R + 31) / 32 * 32;
        b_qweight_ = DeviceTensor({N, K / 2}, 1);
        nb_check(nb200_repack_qweight(qweight.data(), b_qweight_.data(), N, K, use_fp4, stream), "nb200_repack_qweight");
        if (use_fp4) {
            b_wscales_ = DeviceTensor({static_cast<int64_t>(N) * K / 16}, 1);
            nb_check(nb200_repack_wscales_fp4(wscales.data(), b_wscales_.data(), N, K, stream), "nb200_repack_wscales_fp4");
        } else {
            b_wscales_ = DeviceTensor({K / 64, N}, 2);
            nb_check(nb200_repack_wscales_int4(wscales.data(), b_wscales_.data(), N, K, stream), "nb200_repack_wscales_int4");
        }
        b_bias_ = DeviceTensor();
        if (bias.valid()) {
            b_bias_ = DeviceTensor({N}, 4);
            nb_check(nb200_repack_channel_vector(bias.data(), b_bias_.data(), N, dtype, 1, 1.0f, stream), "nb200_repack_channel_vector");
        }
        b_cscale_ = DeviceTensor();
        if (wcscales.numel() > 0) {
            b_cscale_ = DeviceTensor({N}, 4);
            nb_check(nb200_repack_channel_vector(wcscales.data(), b_cscale_.data(), N, dtype, 1, wtscale, stream), "nb200_repack_channel_vector");
        } else if (wtscale != 1.0f) {
            std::vector<float> h(N, wtscale);
            b_cscale_ = DeviceTensor({N}, 4);
            cuda_check(cudaMemcpy(b_cscale_.data(), h.data(), N * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy");
        }
        b_smooth_ = DeviceTensor({K}, 2);
        nb_check(nb200_repack_channel_vector(smooth.data(), b_smooth_.data(), K, dtype, 0, 1.0f, stream), "nb200_repack_channel_vector");
        b_lora_up_ = DeviceTensor({static_cast<int64_t>(N) * Rp}, 2);
        nb_check(nb200_repack_lora_up(lora_up.data(), b_lora_up_.data(), b_cscale_.valid() ? b_cscale_.data_ptr<float>() : nullptr, N, R, dtype, stream),
                 "nb200_repack_lora_up");
        b_lora_down_ = DeviceTensor({2LL * K * Rp}, 2);
        nb_check(nb200_repack_lora_down(lora_down.data(), b_lora_down_.data(), K, R, dtype, stream), "nb200_repack_lora_down");
        b_lora_down_next_ = DeviceTensor({R, K}, 2);
        nb_check(nb200_repack_lora_down_next(lora_down.data(), b_lora_down_next_.data(), K, R, dtype, stream), "nb200_repack_lora_down_next");
        repacked_ = true;
    }

    bool repacked_ = false;
    DeviceTensor b_qweight_, b_wscales_, b_bias_, b_cscale_, b_smooth_, b_lora_up_, b_lora_down_, b_lora_down_next_, workspace_;
};

}  // namespace nunchaku_b200
