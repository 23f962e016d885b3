"""TEST INFRASTRUCTURE -- ctypes binding of ``oracle/_ref/libnunchaku_ref.so``: the UNMODIFIED reference
kernels (src/kernels/zgemm/*, awq/gemv_awq.cu, activation/layernorm/misc kernels) and its C++ module
``GEMM_W4A4`` (src/Linear.cpp), compiled for sm_100a by ``oracle/ref_build/build_ref.sh`` and driven through
``oracle/ref_build/ref_shim.cu``.

Used by: ``tests/golden/make_ref_gpu_golden.py`` (golden vectors), the ``-m gpu`` tests that compare our kernels
with the reference's on the same B200, and ``bench.py``'s ``reference_gpu`` leg.  Never imported by the product.

The same binding drives ``oracle/_ref/libnunchaku_seam.so`` (the reference's host layer linked on top of OUR
kernels) -- pass ``lib="seam"``.

Only the INT4 path can execute on sm_100a: the reference's NVFP4 kernels need ``mma.sync ... block_scale``
(sm_120a) and compile to a trap here (gemm_w4a4.cuh:28-32).
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "ref": os.path.join(_HERE, "_ref", "libnunchaku_ref.so"),
    "seam": os.path.join(_HERE, "_ref", "libnunchaku_seam.so"),
}
_LIBS: dict[str, ctypes.CDLL] = {}

# == reference Tensor::ScalarType (src/Tensor.h:215-226)
_DT = {
    torch.int8: 1, torch.uint8: 1, torch.int16: 2, torch.int32: 3, torch.int64: 4, torch.float16: 5, torch.float32: 6,
    torch.bfloat16: 7, torch.float8_e4m3fn: 8, torch.float8_e5m2: 9,
}


class NrefTensor(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("dtype", ctypes.c_int), ("ndim", ctypes.c_int), ("shape", ctypes.c_int * 5),
                ("on_cpu", ctypes.c_int)]


def available(lib: str = "ref") -> bool:
    return os.path.exists(_PATHS[lib])


def load(lib: str = "ref") -> ctypes.CDLL:
    if lib not in _LIBS:
        if lib == "seam":  # the seam library forwards into the product library
            from nunchaku_b200 import _C  # noqa: F401  (loads libnunchaku_b200.so with RTLD_GLOBAL)
        L = ctypes.CDLL(_PATHS[lib], mode=ctypes.RTLD_GLOBAL if lib == "seam" else ctypes.RTLD_LOCAL)
        L.nref_last_error.restype = ctypes.c_char_p
        L.nref_linear_create.restype = ctypes.c_void_p
        _LIBS[lib] = L
    return _LIBS[lib]


def T(t: torch.Tensor | None) -> ctypes.POINTER(NrefTensor) | None:
    """torch tensor -> nref_tensor* (None -> NULL == the reference's absent Tensor{})."""
    if t is None:
        return None
    assert t.is_contiguous(), "reference ops take contiguous tensors"
    s = NrefTensor()
    s.ptr = t.data_ptr()
    s.dtype = _DT[t.dtype]
    s.ndim = t.dim()
    for i, d in enumerate(t.shape):
        s.shape[i] = d
    s.on_cpu = 0 if t.is_cuda else 1
    return ctypes.pointer(s)


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(L, st: int, what: str) -> None:
    if st != 0:
        raise RuntimeError(f"{what}: {L.nref_last_error().decode()}")


def quantize_w4a4_act_fuse_lora(x, lora_down, smooth=None, fuse_glu=False, fp4=False, pad_size=256, lib="ref"):
    """nunchaku/ops/quantize.py:10-80 semantics (allocates outputs).  ``lora_down``/``smooth`` in checkpoint layout."""
    L = load(lib)
    M, K = x.shape
    if fuse_glu:
        K //= 2
    R = lora_down.shape[1]
    Mp = (M + pad_size - 1) // pad_size * pad_size
    out = torch.empty(Mp, K // 2, dtype=torch.uint8, device=x.device)
    if fp4:
        osc = torch.empty(K // 16, Mp, dtype=torch.float8_e4m3fn, device=x.device)
    else:
        osc = torch.empty(K // 64, Mp, dtype=x.dtype, device=x.device)
    la = torch.empty(Mp, R, dtype=torch.float32, device=x.device)
    _check(L, L.nref_quantize_w4a4_act_fuse_lora(T(x), T(out), T(osc), T(lora_down), T(la), T(smooth), int(fuse_glu), int(fp4), _stream()),
           "quantize_w4a4_act_fuse_lora")
    return out, osc, la


def gemm_w4a4(act, wgt, out=None, qout=None, ascales=None, wscales=None, oscales=None, poolout=None, lora_act_in=None, lora_up=None,
              lora_down=None, lora_act_out=None, norm_q=None, norm_k=None, rotary_emb=None, bias=None, smooth_factor=None,
              out_vk=None, out_linearattn=None, act_unsigned=False, lora_scales=None, fuse_silu=False, fp4=False, alpha=1.0,
              wcscales=None, out_q=None, out_k=None, out_v=None, attn_tokens=0, lib="ref"):
    """29 arguments of nunchaku/ops/gemm.py:12-160 == zgemm.h:8-36."""
    L = load(lib)
    if lora_scales is None:
        rank = lora_up.shape[1] if lora_up is not None else 0
        lora_scales = [1.0] * ((rank + 15) // 16)
    ls = (ctypes.c_float * max(1, len(lora_scales)))(*lora_scales)
    _check(L, L.nref_gemm_w4a4(T(act), T(wgt), T(out), T(qout), T(ascales), T(wscales), T(oscales), T(poolout), T(lora_act_in), T(lora_up),
                               T(lora_down), T(lora_act_out), T(norm_q), T(norm_k), T(rotary_emb), T(bias), T(smooth_factor), T(out_vk),
                               T(out_linearattn), int(act_unsigned), ls, len(lora_scales), int(fuse_silu), int(fp4), ctypes.c_float(alpha),
                               T(wcscales), T(out_q), T(out_k), T(out_v), int(attn_tokens), _stream()), "gemm_w4a4")


def attention_fp16(q, k, v, o, scale, lib="ref"):
    L = load(lib)
    _check(L, L.nref_attention_fp16(T(q), T(k), T(v), T(o), ctypes.c_float(scale), _stream()), "attention_fp16")


def test_rmsnorm_rope(inp, out, norm_q, norm_k, rotary_emb, lib="ref"):
    L = load(lib)
    _check(L, L.nref_test_rmsnorm_rope(T(inp), T(out), T(norm_q), T(norm_k), T(rotary_emb), _stream()), "test_rmsnorm_rope")


def test_pack_qkv(inp, out_q, out_k, out_v, num_tokens, lib="ref"):
    L = load(lib)
    _check(L, L.nref_test_pack_qkv(T(inp), T(out_q), T(out_k), T(out_v), int(num_tokens), _stream()), "test_pack_qkv")


def gemv_awq(x, qweight, scales, zeros, m, n, k, group_size=64, lib="ref"):
    L = load(lib)
    out = torch.empty(m, n, dtype=x.dtype, device=x.device)
    _check(L, L.nref_gemv_awq(T(x), T(qweight), T(scales), T(zeros), m, n, k, group_size, T(out), _stream()), "gemv_awq")
    return out


# ---- glue ------------------------------------------------------------------------------------------------
def glue_activation(kind: str, x, lib="ref"):
    L = load(lib)
    out = torch.empty_like(x)
    _check(L, L.nref_glue_activation({"silu": 0, "gelu": 1}[kind], T(x), T(out), _stream()), kind)
    return out


def glue_layernorm(x, weight, bias, eps, lib="ref"):
    L = load(lib)
    out = torch.empty_like(x)
    _check(L, L.nref_glue_layernorm(T(x), T(weight), T(bias), T(out), ctypes.c_float(eps), _stream()), "layernorm")
    return out


def glue_rms_norm(x, weight, eps, lib="ref"):
    L = load(lib)
    out = torch.empty_like(x)
    _check(L, L.nref_glue_rms_norm(T(x), T(weight), T(out), ctypes.c_float(eps), _stream()), "rms_norm")
    return out


def glue_add(a, b, lib="ref"):
    L = load(lib)
    out = torch.empty_like(a)
    _check(L, L.nref_glue_add(T(a), T(b), T(out), _stream()), "add")
    return out


def glue_mul_add_batch(x, scale, batch_scale, scale_shift, bias, batch_bias, lib="ref"):
    L = load(lib)
    _check(L, L.nref_glue_mul_add_batch(T(x), T(scale), int(batch_scale), ctypes.c_double(scale_shift), T(bias), int(batch_bias), _stream()),
           "mul_add_batch")
    return x


def glue_cast(x, dtype, lib="ref"):
    L = load(lib)
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    _check(L, L.nref_glue_cast(T(x), T(out), _stream()), "cast")
    return out


def glue_split_mod(x, n, lib="ref"):
    L = load(lib)
    shape = list(x.shape)
    shape[-1] //= n
    outs = [torch.empty(shape, dtype=x.dtype, device=x.device) for _ in range(n)]
    arr = (NrefTensor * n)(*[T(o).contents for o in outs])
    _check(L, L.nref_glue_split_mod(T(x), arr, n, _stream()), "split_mod")
    return outs


# ---- class GEMM_W4A4 ------------------------------------------------------------------------------------------
class RefLinear:
    """The reference's C++ module ``GEMM_W4A4`` (src/Linear.h:53-120) on raw checkpoint tensors."""

    def __init__(self, in_features, out_features, bias=True, fp4=False, dtype=torch.bfloat16, lib="ref"):
        self.L = load(lib)
        self.lib = lib
        self.in_features, self.out_features, self.dtype, self.fp4 = in_features, out_features, dtype, fp4
        self.h = self.L.nref_linear_create(in_features, out_features, int(bias), int(fp4), _DT[dtype])
        if not self.h:
            raise RuntimeError(self.L.nref_last_error().decode())
        self.h = ctypes.c_void_p(self.h)

    def load(self, **params):
        for k, v in params.items():
            if v is None:
                continue
            _check(self.L, self.L.nref_linear_load(self.h, k.encode(), T(v.contiguous()), _stream()), f"load {k}")
        torch.cuda.synchronize()
        return self

    @property
    def lora_rank(self):
        return self.L.nref_linear_lora_rank(self.h)

    def set_lora_scales(self, scales):
        arr = (ctypes.c_float * len(scales))(*scales)
        self.L.nref_linear_set_lora_scales(self.h, arr, len(scales))

    def forward(self, x, silu=False):
        out = torch.empty(*x.shape[:-1], self.out_features, dtype=self.dtype, device=x.device)
        _check(self.L, self.L.nref_linear_forward(self.h, T(x), T(out), 2 if silu else 0, _stream()), "GEMM_W4A4::forward")
        return out

    def forward_mlp(self, fc2: "RefLinear", x):
        out = torch.empty(*x.shape[:-1], fc2.out_features, dtype=self.dtype, device=x.device)
        _check(self.L, self.L.nref_linear_forward_mlp(self.h, fc2.h, T(x), T(out), _stream()), "GEMM_W4A4 fused MLP")
        return out

    def forward_qkv(self, x, norm_q, norm_k, rotary_emb, out_q=None, out_k=None, out_v=None, num_tokens=0):
        out = torch.empty(*x.shape[:-1], self.out_features, dtype=self.dtype, device=x.device)
        _check(self.L, self.L.nref_linear_forward_qkv(self.h, T(x), T(out), T(norm_q), T(norm_k), T(rotary_emb), T(out_q), T(out_k), T(out_v),
                                                      int(num_tokens), _stream()), "GEMM_W4A4::forward(qkv)")
        return out

    def __del__(self):
        try:
            torch.cuda.synchronize()
            self.L.nref_linear_destroy(self.h)
        except Exception:
            pass
