// Drop-in definitions of the reference's elementwise / norm helpers on top of
// libnunchaku_b200.so, with the reference's own signatures:
//     silu, gelu_new                      src/kernels/activation_kernels.h      (callers: src/activation.cpp:4-14)
//     layernorm_general, rms_norm         src/kernels/layernorm_kernels.h       (callers: src/layernorm.cpp:14-24)
//     kernels::add / mul_add / mul_add_batch / cast / split_mod<N>   src/kernels/misc_kernels.h:8-25
// Replaces src/kernels/{activation,layernorm,misc}_kernels.cu in the reference build; src/activation.cpp,
// src/layernorm.cpp, src/Module.cpp, src/Linear.cpp, src/FluxModel.cpp compile and link unchanged
// (oracle/ref_build/build_ref.sh links exactly that into oracle/_ref/libnunchaku_seam.so).
#include <array>

#include "Tensor.h"
#include "common.h"
#include "kernels/activation_kernels.h"
#include "kernels/layernorm_kernels.h"
#include "kernels/misc_kernels.h"
#include "nunchaku_b200.h"

namespace {
int code(const Tensor &t) {
    switch (t.scalar_type()) {
        case Tensor::FP16: return NB200_FP16;
        case Tensor::BF16: return NB200_BF16;
        case Tensor::FP32: return NB200_FP32;
        default: throw std::invalid_argument("nunchaku_b200 glue: fp16 / bf16 / fp32 tensor expected");
    }
}
void check(int st, const char *what) {
    if (st != NB200_OK) throw std::runtime_error(std::string(what) + ": " + nb200_last_error());
}
void *stream() { return getCurrentCUDAStream(); }
void require_contiguous(const Tensor &t, const char *what) {
    if (!t.is_contiguous()) throw std::invalid_argument(std::string(what) + ": contiguous tensor expected");
}
}  // namespace

void silu(Tensor &out, Tensor &input) {
    check(nb200_activation(NB200_ACT_SILU, code(input), input.data_ptr(), out.data_ptr(), (long long)input.numel(), stream()), "nb200_activation");
}
void gelu_new(Tensor &out, Tensor &input) {
    check(nb200_activation(NB200_ACT_GELU, code(input), input.data_ptr(), out.data_ptr(), (long long)input.numel(), stream()), "nb200_activation");
}

void layernorm_general(Tensor out, Tensor input, Tensor weight, Tensor bias, float epsilon) {
    const int hidden = input.shape[-1];
    check(nb200_layernorm(code(input), input.data_ptr(), weight.valid() ? weight.data_ptr() : nullptr, bias.valid() ? bias.data_ptr() : nullptr,
                          out.data_ptr(), (long long)(input.numel() / hidden), hidden, epsilon, stream()),
          "nb200_layernorm");
}
void rms_norm(Tensor &out, Tensor &input, Tensor &weight, float epsilon, bool use_quant) {
    if (use_quant) throw std::runtime_error("rms_norm(use_quant): W8A8 LLM path, out of scope");
    const int hidden = input.shape[-1];
    check(nb200_rms_norm(code(input), input.data_ptr(), weight.data_ptr(), out.data_ptr(), (long long)(input.numel() / hidden), hidden, epsilon,
                         stream()),
          "nb200_rms_norm");
}
// LLM-serving leftovers layernorm.cpp references (RMSNormGeneral): never instantiated by FluxModel / SanaModel
void rms_norm_general(Tensor &, Tensor &, Tensor &, Tensor &, float, bool) { throw std::runtime_error("rms_norm_general: out of scope"); }
void rms_norm_general_fuse_sum(Tensor &, Tensor &, Tensor &, Tensor &, Tensor &, float, bool) {
    throw std::runtime_error("rms_norm_general_fuse_sum: out of scope");
}

namespace nunchaku::kernels {

Tensor add(Tensor a, Tensor b) {
    require_contiguous(a, "add");
    require_contiguous(b, "add");
    Tensor out = Tensor::empty_like(a);
    check(nb200_add(code(a), a.data_ptr(), b.data_ptr(), out.data_ptr(), (long long)a.numel(), stream()), "nb200_add");
    return out;
}

void mul_add(Tensor x, Tensor scale, Tensor bias) {
    require_contiguous(x, "mul_add");
    check(nb200_mul_add_batch(code(x), x.data_ptr(), scale.valid() ? scale.data_ptr() : nullptr, bias.data_ptr(), 0.0f, 1, (long long)x.numel(),
                              scale.valid() ? (long long)scale.numel() : 1, (long long)bias.numel(), 0, 0, 0, stream()),
          "nb200_mul_add_batch");
}

void mul_add_batch(Tensor x, Tensor scale, bool batch_scale, double scale_shift, Tensor bias, bool batch_bias) {
    const int batch = x.shape[0];
    const long long numel = (long long)x.numel() / batch;
    const long long numel_scale = scale.valid() ? (long long)scale.numel() / (batch_scale ? batch : 1) : 1;
    const long long numel_bias = (long long)bias.numel() / (batch_bias ? batch : 1);
    check(nb200_mul_add_batch(code(x), x.data_ptr(), scale.valid() ? scale.data_ptr() : nullptr, bias.data_ptr(), (float)scale_shift, batch, numel,
                              numel_scale, numel_bias, (long long)x.stride(0), (scale.valid() && batch_scale) ? (long long)scale.stride(0) : 0,
                              batch_bias ? (long long)bias.stride(0) : 0, stream()),
          "nb200_mul_add_batch");
}

void cast(Tensor input, Tensor output) {
    require_contiguous(input, "cast");
    require_contiguous(output, "cast");
    check(nb200_cast(code(input), input.data_ptr(), code(output), output.data_ptr(), (long long)input.numel(), stream()), "nb200_cast");
}

template <size_t N>
std::array<Tensor, N> split_mod(Tensor input) {
    require_contiguous(input, "split_mod");
    auto shape = TensorShape(input.shape.dataExtent);
    shape[-1] /= int(N);
    std::array<Tensor, N> outs;
    void *ptrs[N];
    for (size_t i = 0; i < N; i++) {
        outs[i] = Tensor::empty(shape, input.scalar_type(), input.device());
        ptrs[i] = outs[i].data_ptr();
    }
    check(nb200_split_mod(code(input), input.data_ptr(), ptrs, int(N), (long long)input.numel(), stream()), "nb200_split_mod");
    return outs;
}
template std::array<Tensor, 2> split_mod<2>(Tensor);
template std::array<Tensor, 3> split_mod<3>(Tensor);
template std::array<Tensor, 4> split_mod<4>(Tensor);
template std::array<Tensor, 5> split_mod<5>(Tensor);
template std::array<Tensor, 6> split_mod<6>(Tensor);

}  // namespace nunchaku::kernels
