"""CPU: host-side contract of the operator mirror -- what is rejected, and how (no GPU needed).

The reference aborts on precondition failures (`assert` with -UNDEBUG); the mirror raises instead, and it
never falls back to a CPU computation: CPU tensors are an error, not a slow path."""
import pytest
import torch

from nunchaku_b200.models.linear import SVDQW4A4Linear
from nunchaku_b200.ops import glue
from nunchaku_b200.ops.fused import _fuse_fc1
from nunchaku_b200.ops.gemm import linearattn_vk_mul_q, svdq_gemm_w4a4_cuda
from nunchaku_b200.ops.quantize import svdq_quantize_w4a4_act_fuse_lora_cuda
from nunchaku_b200.utils import ceil_divide, get_precision, torch_dtype_code


def test_gemm_rejects_missing_and_cpu_operands():
    with pytest.raises(ValueError):
        svdq_gemm_w4a4_cuda(act=None, wgt=None, ascales=None, wscales=None)
    act = torch.zeros(256, 64, dtype=torch.uint8)
    wgt = torch.zeros(128, 64, dtype=torch.int8)
    sc = torch.zeros(2, 256, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        svdq_gemm_w4a4_cuda(act=act, wgt=wgt, out=torch.zeros(256, 128, dtype=torch.bfloat16), ascales=sc, wscales=sc)


def test_quantize_rejects_bad_inputs():
    x = torch.zeros(4, 128, dtype=torch.bfloat16)
    ld = torch.zeros(128, 32, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        svdq_quantize_w4a4_act_fuse_lora_cuda(x.view(1, 4, 128), lora_down=ld)       # 2-D only
    with pytest.raises(ValueError):
        svdq_quantize_w4a4_act_fuse_lora_cuda(x, lora_down=None)                      # lora_down is required
    with pytest.raises(RuntimeError, match="no CPU path"):
        svdq_quantize_w4a4_act_fuse_lora_cuda(x, lora_down=ld)


def test_glue_rejects_cpu_tensors_and_bad_dtypes():
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    for fn in (glue.silu, glue.gelu_new):
        with pytest.raises(RuntimeError, match="no CPU path"):
            fn(x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        glue.add(x, x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        glue.layernorm(x)
    with pytest.raises(RuntimeError):
        linearattn_vk_mul_q(x, torch.zeros(1, 2, 33, 32))


def test_module_mirror_surface_and_cpu_forward():
    """Same constructor keywords, parameter names, shapes and dtypes as nunchaku/models/linear.py:13-120."""
    m = SVDQW4A4Linear(256, 384, rank=32, bias=True, precision="int4", torch_dtype=torch.bfloat16, device="cpu")
    sd = m.state_dict()
    assert sd["qweight"].shape == (384, 128) and sd["qweight"].dtype == torch.int8
    assert sd["wscales"].shape == (256 // 64, 384) and sd["wscales"].dtype == torch.bfloat16
    assert sd["proj_down"].shape == (256, 32) and sd["proj_up"].shape == (384, 32)
    assert sd["smooth_factor"].shape == (256,) and sd["bias"].shape == (384,)
    f = SVDQW4A4Linear(256, 384, rank=16, bias=False, precision="nvfp4", torch_dtype=torch.float16, device="cpu")
    fs = f.state_dict()
    assert fs["wscales"].shape == (256 // 16, 384) and fs["wscales"].dtype == torch.float8_e4m3fn
    assert fs["wcscales"].shape == (384,) and f.wtscale == 1.0 and f.bias is None
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 256, dtype=torch.bfloat16))                             # CPU forward is an error, not a fallback


def test_small_helpers():
    assert ceil_divide(300, 256) == 2 and ceil_divide(256, 256) == 1
    assert torch_dtype_code(torch.float16) == 0 and torch_dtype_code(torch.bfloat16) == 1
    with pytest.raises(TypeError):
        torch_dtype_code(torch.float32)
    assert get_precision("auto") == "nvfp4" and get_precision("fp4") == "nvfp4" and get_precision("int4") == "int4"
    assert get_precision("auto", pretrained_model_name_or_path="svdq-int4_r32-flux.1-schnell") == "int4"

    class _L:
        precision = "nvfp4"
        out_features = 12288

    assert _fuse_fc1(_L, 256) is True and _fuse_fc1(_L, 4352) is False               # small M fuses, large M splits


def test_on_device_of_passes_through_and_preserves_signature():
    import inspect

    from nunchaku_b200.utils import on_device_of

    calls = []

    @on_device_of("x")
    def op(x, y=2):
        calls.append((x, y))
        return "ok"

    assert op(torch.zeros(1)) == "ok" and op(x=None, y=3) == "ok" and op() if False else True   # CPU / missing tensors fall through
    assert list(inspect.signature(op).parameters) == ["x", "y"] and op.__name__ == "op"
    # the mirrored op keeps the reference's 29 positional parameters (+ the keyword-only extension)
    params = inspect.signature(svdq_gemm_w4a4_cuda).parameters
    assert list(params)[:5] == ["act", "wgt", "out", "qout", "ascales"] and list(params)[28] == "attn_tokens"
    assert params["fuse_gelu"].kind is inspect.Parameter.KEYWORD_ONLY


def test_round2_ops_reject_cpu_tensors_and_keep_reference_signatures():
    """attention / AWQ GEMV / depthwise conv / AdaLN glue / graph capture: errors on CPU, never a fallback; reference argument order kept."""
    import inspect

    from nunchaku_b200.graph import GraphedStep
    from nunchaku_b200.ops.attention import attention_fp16
    from nunchaku_b200.ops.dwconv import DWCONV, dwconv_f16
    from nunchaku_b200.ops.gemv import AWQW4A16Linear, awq_gemv_w4a16_cuda

    q = torch.zeros(1, 2, 128, 128, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        attention_fp16(q, q, q, torch.zeros(1, 128, 256, dtype=torch.float16), 0.1)
    assert list(inspect.signature(attention_fp16).parameters) == ["q", "k", "v", "o", "scale"]          # _C.ops.attention_fp16 (csrc/ops.h)
    with pytest.raises(RuntimeError, match="no CPU path"):
        awq_gemv_w4a16_cuda(torch.zeros(1, 256, dtype=torch.bfloat16), torch.zeros(2, 128, dtype=torch.int32), torch.zeros(4, 8, dtype=torch.bfloat16),
                            torch.zeros(4, 8, dtype=torch.bfloat16), 1, 8, 256)
    p = inspect.signature(awq_gemv_w4a16_cuda).parameters
    assert list(p)[:8] == ["in_feats", "kernel", "scaling_factors", "zeros", "m", "n", "k", "group_size"]  # nunchaku/ops/gemv.py:10-58
    assert p["bias"].kind is inspect.Parameter.KEYWORD_ONLY and p["fuse_silu"].kind is inspect.Parameter.KEYWORD_ONLY
    m = AWQW4A16Linear(256, 64, torch_dtype=torch.bfloat16)
    assert m.qweight.shape == (16, 128) and m.qweight.dtype == torch.int32 and m.wscales.shape == (4, 64) and m.bias.shape == (64,)
    x = torch.zeros(1, 4, 4, 16, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        dwconv_f16(x, torch.zeros(16, 3, 3, 1, dtype=torch.bfloat16))
    assert list(inspect.signature(dwconv_f16).parameters) == ["input", "weight", "out", "bias"]          # src/kernels/dwconv.h:9
    d = DWCONV(16, True, torch.float16)
    assert d.weight.shape == (16, 3, 3, 1) and d.bias.shape == (16,)                                     # src/Linear.cpp:541-547
    with pytest.raises(RuntimeError, match="no CPU path"):
        glue.layernorm_mod(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(64, dtype=torch.bfloat16), torch.zeros(64, dtype=torch.bfloat16))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            GraphedStep(lambda: None)
    assert inspect.signature(svdq_gemm_w4a4_cuda).parameters["qkv_scratch"].kind is inspect.Parameter.KEYWORD_ONLY


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under nunchaku_b200/ may import it (or name its libraries), and bench.py only inside its CPU /
    reference comparison legs -- the timed GPU path (StackRunner / FullRunner / main's timed regions) stays clear of it."""
    import ast
    import glob
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in glob.glob(os.path.join(root, "nunchaku_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        tree = ast.parse(src)
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), f"{path} imports the oracle"
        assert "oracle/_ref" not in src and "libsvdq_ref" not in src, path
    for path in glob.glob(os.path.join(root, "nunchaku_b200", "csrc", "*.cu")) + glob.glob(os.path.join(root, "nunchaku_b200", "csrc", "*.cuh")):
        assert "oracle/" not in open(path).read().replace("oracle/ref_build", ""), path   # (comments may name the build recipe of the seams)
    # bench.py: the oracle appears only in the functions of the comparison legs
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    allowed = {"reference_gpu_leg", "CpuSample", "__init__"}
    for fn in [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.ClassDef))]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                assert fn.name in allowed, f"bench.py: {fn.name} imports the oracle"
