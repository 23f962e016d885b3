"""ctypes binding of the C ABI declared in include/nunchaku_b200.h.

This module plays the role of the reference's pybind module ``nunchaku._C`` for the SVDQuant
path (nunchaku/csrc/pybind.cpp:11-124): ``from nunchaku_b200._C import lib``.  There is NO
fallback: if the shared library is missing or fails to load the import raises, and every op
raises ``RuntimeError`` (message from ``nb200_last_error``) on a non-zero status.
"""
from __future__ import annotations

import ctypes
import os

from ._build import LIB_PATH

NB200_MAX_LORA_SCALES = 64
NB200_FP16, NB200_BF16 = 0, 1
NB200_ACT_NONE, NB200_ACT_SILU, NB200_ACT_GELU = 0, 1, 2

c_void_p, c_int, c_float, c_longlong = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong


class QuantizeArgs(ctypes.Structure):
    _fields_ = [
        ("input", c_void_p),
        ("output", c_void_p),
        ("oscales", c_void_p),
        ("lora_down", c_void_p),
        ("lora_act_out", c_void_p),
        ("smooth", c_void_p),
        ("M", c_int),
        ("Mp", c_int),
        ("K", c_int),
        ("R", c_int),
        ("dtype", c_int),
        ("fuse_glu", c_int),
        ("fp4", c_int),
        ("workspace", c_void_p),
        ("workspace_bytes", ctypes.c_longlong),
        ("act_unsigned_shift", c_int),
    ]


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("act", c_void_p),
        ("wgt", c_void_p),
        ("ascales", c_void_p),
        ("wscales", c_void_p),
        ("bias", c_void_p),
        ("cscale", c_void_p),
        ("lora_act_in", c_void_p),
        ("lora_up", c_void_p),
        ("out", c_void_p),
        ("qout", c_void_p),
        ("oscales", c_void_p),
        ("smooth_next", c_void_p),
        ("lora_down_next", c_void_p),
        ("lora_act_out", c_void_p),
        ("norm_q", c_void_p),
        ("norm_k", c_void_p),
        ("rotary_emb", c_void_p),
        ("Mp", c_int),
        ("N", c_int),
        ("K", c_int),
        ("M_out", c_int),
        ("N_out", c_int),
        ("R_up", c_int),
        ("R_down", c_int),
        ("dtype", c_int),
        ("fp4", c_int),
        ("act_unsigned", c_int),
        ("mid_act", c_int),
        ("lora_scales", c_float * NB200_MAX_LORA_SCALES),
        ("block_n", c_int),
        ("num_sms", c_int),
        ("prof", c_void_p),
        ("out_q", c_void_p),
        ("out_k", c_void_p),
        ("out_v", c_void_p),
        ("stride_head_q", c_longlong),
        ("stride_head_k", c_longlong),
        ("stride_head_v", c_longlong),
        ("attn_tokens", c_int),
        ("workspace", c_void_p),
        ("workspace_bytes", c_longlong),
        ("out_vk", c_void_p),
        ("vk_tokens", c_int),
    ]


# every symbol include/nunchaku_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "nb200_abi_version": (c_int, []),
    "nb200_last_error": (ctypes.c_char_p, []),
    "nb200_check_device": (c_int, []),
    "nb200_last_launch_count": (c_int, []),
    "nb200_repack_qweight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "nb200_repack_wscales_int4": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "nb200_repack_wscales_fp4": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "nb200_repack_channel_vector": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "nb200_repack_lora_up": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "nb200_repack_lora_down": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "nb200_repack_lora_down_next": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "nb200_quantize_workspace_bytes": (ctypes.c_longlong, [c_int, c_int]),
    "nb200_quantize_w4a4_act_fuse_lora": (c_int, [ctypes.POINTER(QuantizeArgs), c_void_p]),
    "nb200_gemm_w4a4": (c_int, [ctypes.POINTER(GemmArgs), c_void_p]),
    "nb200_activation": (c_int, [c_int, c_int, c_void_p, c_void_p, c_longlong, c_void_p]),
    "nb200_layernorm": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_void_p]),
    "nb200_layernorm_mod": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_longlong, c_int, c_float, c_void_p]),
    "nb200_rms_norm": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_void_p]),
    "nb200_add": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "nb200_mul_add_batch": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_float, c_int, c_longlong, c_longlong, c_longlong,
                                    c_longlong, c_longlong, c_longlong, c_void_p]),
    "nb200_split_mod": (c_int, [c_int, c_void_p, ctypes.POINTER(c_void_p), c_int, c_longlong, c_void_p]),
    "nb200_cast": (c_int, [c_int, c_void_p, c_int, c_void_p, c_longlong, c_void_p]),
    "nb200_litela_vk": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "nb200_linearattn_vk_mul_q": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "nb200_attention_fp16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "nb200_gemv_awq": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "nb200_gemm_workspace_bytes": (c_longlong, [c_int, c_int]),
    "nb200_dwconv3x3": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "nb200_gemv_awq_fused": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
}


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"nunchaku_b200: native library not built ({LIB_PATH}). Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU/PyTorch fallback."
        )
    cdll = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(cdll, name)  # AttributeError if the .so is stale / incomplete
        fn.restype = restype
        fn.argtypes = argtypes
    return cdll


lib = _load()


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib.nb200_last_error()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


assert lib.nb200_abi_version() == 2, "nunchaku_b200: stale shared library (ABI version mismatch); rebuild"
