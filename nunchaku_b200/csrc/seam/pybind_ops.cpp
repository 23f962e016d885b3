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
// `_C.ops.attention_fp16` and `_C.ops.gemv_awq` resolve to our kernels as well (zgemm_b200.cpp, awq_b200.cpp).  The remaining entries of
// ops.h (gemm_awq, test_*) are registered too, because ops.h defines them, and raise.
#include <torch/extension.h>

#include "ops.h"   // the reference's nunchaku/csrc/ops.h

namespace nunchaku::kernels {
void b200_invalidate_all();
void b200_set_identity(const void *device_ptr, uint64_t token);
void b200_clear_identities();
// reference-internal test hooks that ops.h references: not provided by the B200 library
void test_rmsnorm_rope(Tensor, Tensor, Tensor, Tensor, Tensor) { throw std::runtime_error("test_rmsnorm_rope: reference-internal test hook, not provided"); }
void test_pack_qkv(Tensor, Tensor, Tensor, Tensor, int) { throw std::runtime_error("test_pack_qkv: reference-internal test hook, not provided"); }
}  // namespace nunchaku::kernels
Tensor awq_gemm_forward_cuda(Tensor, Tensor, Tensor, Tensor) { throw std::runtime_error("awq_gemm_forward_cuda: not provided by nunchaku_b200"); }

namespace {

// Identity of a weight tensor for the converted-weight cache (zgemm_b200.cpp): torch's allocator reuses device addresses, so the
// key is the live StorageImpl (a weak reference keeps the object -- not the memory -- from being recycled under the same
// address) plus the tensor's version counter (in-place updates: LoRA merges, load_state_dict copy_).
struct StorageId {
    c10::weak_intrusive_ptr<c10::StorageImpl> weak;
    uint64_t id;
};
std::mutex g_id_mu;
std::unordered_map<const c10::StorageImpl *, StorageId> g_ids;
uint64_t g_next_id = 1;

void tag(const std::optional<torch::Tensor> &t) {
    if (!t.has_value() || !t->defined() || !t->is_cuda()) return;
    c10::StorageImpl *impl = t->storage().unsafeGetStorageImpl();
    uint64_t id;
    {
        std::lock_guard<std::mutex> lock(g_id_mu);
        auto it = g_ids.find(impl);
        if (it != g_ids.end() && !it->second.weak.expired()) {
            id = it->second.id;
        } else {
            if (g_ids.size() > 4096)   // drop records of storages that died
                for (auto i = g_ids.begin(); i != g_ids.end();) i = i->second.weak.expired() ? g_ids.erase(i) : std::next(i);
            id = g_next_id++;
            auto strong = c10::intrusive_ptr<c10::StorageImpl>::reclaim_copy(impl);
            g_ids.insert_or_assign(impl, StorageId{c10::weak_intrusive_ptr<c10::StorageImpl>(strong), id});
        }
    }
    nunchaku::kernels::b200_set_identity(t->data_ptr(), (id << 24) ^ uint64_t(t->_version()));
}
struct IdentityScope {
    ~IdentityScope() { nunchaku::kernels::b200_clear_identities(); }
};
using OT = std::optional<torch::Tensor>;

void gemm_w4a4_tagged(OT act, OT wgt, OT out, OT qout, OT ascales, OT wscales, OT oscales, OT poolout, OT lora_act_in, OT lora_up, OT lora_down,
                      OT lora_act_out, OT norm_q, OT norm_k, OT rotary_emb, OT bias, OT smooth_factor, OT out_vk, OT out_linearattn, bool act_unsigned,
                      std::vector<float> lora_scales, bool fuse_silu, bool fp4, float alpha, OT wcscales, OT out_q, OT out_k, OT out_v, int attn_tokens) {
    IdentityScope scope;
    for (const OT *t : {&wgt, &wscales, &lora_up, &lora_down, &bias, &smooth_factor, &wcscales}) tag(*t);
    nunchaku::ops::gemm_w4a4(act, wgt, out, qout, ascales, wscales, oscales, poolout, lora_act_in, lora_up, lora_down, lora_act_out, norm_q, norm_k, rotary_emb,
                             bias, smooth_factor, out_vk, out_linearattn, act_unsigned, lora_scales, fuse_silu, fp4, alpha, wcscales, out_q, out_k, out_v,
                             attn_tokens);
}
void quantize_tagged(OT input, OT output, OT oscales, OT lora_down, OT lora_act_out, OT smooth, bool fuse_glu, bool fp4) {
    IdentityScope scope;
    tag(lora_down);
    tag(smooth);
    nunchaku::ops::quantize_w4a4_act_fuse_lora(input, output, oscales, lora_down, lora_act_out, smooth, fuse_glu, fp4);
}

}  // namespace

PYBIND11_MODULE(_C, m) {
    m.doc() = "nunchaku._C surface of the SVDQuant hot path on libnunchaku_b200.so";
    m.def_submodule("ops")
        .def("gemm_w4a4", gemm_w4a4_tagged)
        .def("quantize_w4a4_act_fuse_lora", quantize_tagged)
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
