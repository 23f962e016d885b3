// `nunchaku._C` for B200: the pybind surface Route A / Route B callers bind to (reference nunchaku/csrc/pybind.cpp:11-124),
// restricted to the SVDQuant hot path:
//
//     _C.ops.gemm_w4a4(29 positional args, Optional[Tensor])          nunchaku/csrc/ops.h:10-81
//     _C.ops.quantize_w4a4_act_fuse_lora(8 args)                      nunchaku/csrc/ops.h:83-112
//     _C.utils.{set_log_level, set_cuda_stack_limit, disable_memory_auto_release, trim_memory, set_faster_i2f_mode}
//
// The binding functions themselves are the REFERENCE'S OWN (`#include "ops.h"` below, compiled where it lies, together with
// its src/interop/torch.cpp): torch tensors become the reference's `Tensor` views, the current torch stream is pushed by its
// `TorchOpContext`, and `nunchaku::kernels::gemm_w4a4 / quantize_w4a4_act_fuse_lora` resolve to OUR definitions
// (zgemm_b200.cpp -> libnunchaku_b200.so).  Built by oracle/ref_build/build_ref.sh into oracle/_ref/pyseam/_C.so;
// tests/test_gpu_seam_pybind.py imports it and checks it against the Python operator layer bit for bit.
//
// Out-of-scope entries of ops.h (attention_fp16, gemm_awq, test_*) are registered too, because ops.h defines them, and raise.
#include <torch/extension.h>

#include "ops.h"   // the reference's nunchaku/csrc/ops.h

namespace nunchaku::kernels {
void b200_invalidate_all();
// out-of-scope kernels that ops.h references: not provided by the B200 library (SURVEY section 8f rows N1 / N2 are "next")
void attention_fp16(Tensor, Tensor, Tensor, Tensor, float) { throw std::runtime_error("attention_fp16: not provided by nunchaku_b200 (SURVEY N1)"); }
void test_rmsnorm_rope(Tensor, Tensor, Tensor, Tensor, Tensor) { throw std::runtime_error("test_rmsnorm_rope: reference-internal test hook, not provided"); }
void test_pack_qkv(Tensor, Tensor, Tensor, Tensor, int) { throw std::runtime_error("test_pack_qkv: reference-internal test hook, not provided"); }
}  // namespace nunchaku::kernels
Tensor awq_gemm_forward_cuda(Tensor, Tensor, Tensor, Tensor) { throw std::runtime_error("awq_gemm_forward_cuda: not provided by nunchaku_b200"); }

PYBIND11_MODULE(_C, m) {
    m.doc() = "nunchaku._C surface of the SVDQuant hot path on libnunchaku_b200.so";
    m.def_submodule("ops")
        .def("gemm_w4a4", nunchaku::ops::gemm_w4a4)
        .def("quantize_w4a4_act_fuse_lora", nunchaku::ops::quantize_w4a4_act_fuse_lora)
        .def("attention_fp16", nunchaku::ops::attention_fp16)
        .def("gemv_awq", nunchaku::ops::gemv_awq)
        .def("gemm_awq", nunchaku::ops::gemm_awq)
        .def("test_rmsnorm_rope", nunchaku::ops::test_rmsnorm_rope)
        .def("test_pack_qkv", nunchaku::ops::test_pack_qkv)
        // addition: drop the converted-weight cache after parameters were changed in place (zgemm_b200.cpp)
        .def("b200_invalidate_all", [] { nunchaku::kernels::b200_invalidate_all(); });
    m.def_submodule("utils")
        .def("set_log_level", [](const std::string &) {})
        .def("set_cuda_stack_limit", [](int64_t) {})
        .def("disable_memory_auto_release", [] {})
        .def("trim_memory", [] {})
        .def("set_faster_i2f_mode", [](const std::string &) {});
}
