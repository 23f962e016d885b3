"""The `nunchaku._C.ops` pybind surface on B200 (SURVEY section 8b "Python seam"): oracle/_ref/pyseam/_C.so is the REFERENCE'S OWN
binding code (nunchaku/csrc/ops.h + src/interop/torch.cpp, compiled where they lie by oracle/ref_build/build_ref.sh) linked on top of
our definitions of the zgemm.h functions (nunchaku_b200/csrc/seam/zgemm_b200.cpp -> libnunchaku_b200.so).  Route A / Route B
callers and `QuantizedGEMM` bind exactly here: 29 positional Optional[Tensor] arguments for gemm_w4a4, 8 for the quantizer.
Checked bit for bit against this package's Python operator layer on the same checkpoint tensors."""
import importlib.util
import os

import pytest
import torch

from gpu_util import ref_layout_params
from oracle import svdq as O

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "pyseam", "_C.so")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/pyseam/_C.so not built")]


@pytest.fixture(scope="module")
def C():
    from nunchaku_b200 import _C as _lib  # noqa: F401  (libnunchaku_b200.so first)

    spec = importlib.util.spec_from_file_location("_C", SO)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
def test_pybind_ops_match_python_operator_layer(C, fp4, hT):
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda
    from nunchaku_b200.ops.quantize import svdq_quantize_w4a4_act_fuse_lora_cuda

    N, K, R, M = 512, 384, 32, 300
    layer = O.make_synthetic_layer(N, K, R, fp4=fp4, hT=hT, seed=501)
    p = ref_layout_params(layer)
    x = O.make_activations(M, K, hT, seed=502, smooth=layer.smooth).cuda()
    Mp = 512
    # the reference's Python wrappers allocate the outputs and pass everything positionally (nunchaku/ops/quantize.py:68-79, gemm.py:130-160)
    act = torch.empty(Mp, K // 2, dtype=torch.uint8, device="cuda")
    asc = torch.empty(K // 16, Mp, dtype=torch.float8_e4m3fn, device="cuda") if fp4 else torch.empty(K // 64, Mp, dtype=hT, device="cuda")
    la = torch.empty(Mp, R, dtype=torch.float32, device="cuda")
    C.ops.quantize_w4a4_act_fuse_lora(x, act, asc, p["proj_down"], la, p["smooth"], False, fp4)
    out = torch.full((M, N), float("nan"), dtype=hT, device="cuda")
    C.ops.gemm_w4a4(act, p["qweight"], out, None, asc, p["wscales"], None, None, la, p["proj_up"], None, None, None, None, None, p["bias"], None, None,
                    None, False, [1.0, 1.0], False, fp4, float(layer.alpha), p["wcscales"], None, None, None, 0)
    q2, s2, la2 = svdq_quantize_w4a4_act_fuse_lora_cuda(x, lora_down=p["proj_down"], smooth=p["smooth"], fp4=fp4)
    out2 = torch.empty(M, N, dtype=hT, device="cuda")
    svdq_gemm_w4a4_cuda(act=q2, wgt=p["qweight"], out=out2, ascales=s2, wscales=p["wscales"], lora_act_in=la2, lora_up=p["proj_up"], bias=p["bias"],
                        fp4=fp4, alpha=layer.alpha, wcscales=p["wcscales"])
    torch.cuda.synchronize()
    assert torch.equal(act, q2) and torch.equal(asc.view(torch.uint8), s2.view(torch.uint8)) and torch.equal(la, la2)
    assert torch.equal(out, out2)
    y_ref = O.svdq_linear_forward(layer, x.cpu(), mode="ref")
    assert O.rel_fro(out.cpu(), y_ref) <= 1.5e-2


def test_pybind_cache_survives_address_reuse_and_in_place_updates(C):
    """torch's allocator hands a freed weight's device address to the next weight of the same size, and LoRA merges / load_state_dict
    update parameters in place: the converted-weight cache behind `_C.ops` is keyed on the live storage + version counter, not on the
    address alone (csrc/seam/pybind_ops.cpp)"""
    N, K, R, M, Mp = 256, 256, 32, 256, 256
    hT = torch.bfloat16
    outs = []
    for seed in (601, 602):
        layer = O.make_synthetic_layer(N, K, R, fp4=True, hT=hT, seed=seed)
        p = ref_layout_params(layer)
        x = O.make_activations(M, K, hT, seed=77, smooth=layer.smooth).cuda()
        act = torch.empty(Mp, K // 2, dtype=torch.uint8, device="cuda")
        asc = torch.empty(K // 16, Mp, dtype=torch.float8_e4m3fn, device="cuda")
        la = torch.empty(Mp, R, dtype=torch.float32, device="cuda")

        def run():
            C.ops.quantize_w4a4_act_fuse_lora(x, act, asc, p["proj_down"], la, p["smooth"], False, True)
            out = torch.empty(M, N, dtype=hT, device="cuda")
            C.ops.gemm_w4a4(act, p["qweight"], out, None, asc, p["wscales"], None, None, la, p["proj_up"], None, None, None, None, None, p["bias"], None,
                            None, None, False, [1.0, 1.0], False, True, float(layer.alpha), p["wcscales"], None, None, None, 0)
            return out

        out = run()
        assert O.rel_fro(out.cpu(), O.svdq_linear_forward(layer, x.cpu(), mode="ref")) <= 1.5e-2, seed
        ptrs = {k: v.data_ptr() for k, v in p.items() if v is not None}
        outs.append(out.clone())
        # in-place update of a parameter (same storage, same address): the next call must see it
        p["bias"].add_(1.0)
        out_b = run()
        assert O.rel_fro((out_b.float() - out.float()).cpu(), torch.ones(M, N)) <= 2e-2
        del p, layer, out, out_b   # free: the next iteration's parameters are likely to land on the same addresses
    assert not torch.equal(outs[0], outs[1]) and ptrs


def test_pybind_utils_and_out_of_scope_entries(C):
    C.utils.set_log_level("info")
    C.utils.disable_memory_auto_release()
    C.utils.trim_memory()
    with pytest.raises(RuntimeError, match="not provided"):
        t = torch.zeros(1, 256, 128, dtype=torch.float16, device="cuda")
        C.ops.test_rmsnorm_rope(t, t, torch.ones(128, dtype=torch.float16, device="cuda"), torch.ones(128, dtype=torch.float16, device="cuda"),
                                torch.zeros(1, 256, 128, dtype=torch.float32, device="cuda"))


def test_pybind_attention_fp16_is_our_kernel(C):
    """`_C.ops.attention_fp16(q, k, v, o, scale)` (ops.h) on csrc/attention.cu: same result as the Python operator"""
    from nunchaku_b200.ops.attention import attention_fp16
    from oracle import attention as AT

    g = torch.Generator().manual_seed(4)
    qkv = (torch.randn(200, 3 * 2 * 128, generator=g) * 0.5).to(torch.float16)
    q, k, v = (t.cuda() for t in AT.pack_qkv_rowmajor(qkv, heads=2, tokens_pad=256))
    o1 = torch.empty(1, 256, 256, dtype=torch.float16, device="cuda")
    o2 = torch.empty_like(o1)
    C.ops.attention_fp16(q, k, v, o1, 128 ** -0.5)
    attention_fp16(q, k, v, o2, 128 ** -0.5)
    torch.cuda.synchronize()
    assert torch.equal(o1[:, :200], o2[:, :200])
    assert O.rel_fro(o1[:, :200].cpu(), AT.attention_fp16(q.cpu(), k.cpu(), v.cpu(), 128 ** -0.5)[:, :200]) <= 6e-4
