"""Attention of the FLUX blocks on B200 -- the reference's operator signature (``nunchaku._C.ops.attention_fp16(q, k, v, o, scale)``,
nunchaku/csrc/ops.h; NunchakuFP16AttnProcessor, nunchaku/models/attention_processors/flux.py) over ``nb200_attention_fp16``.

q / k / v are what the QKV projection's PackQKV epilogue wrote (``fused_qkv_norm_rottary(..., output=(q, k, v), attn_tokens=T)``): fp16
``[B, H, T_pad, 128]``, row-major inside a head, K's pad rows NaN (the key mask), Q's and V's pad rows 0."""
from __future__ import annotations

import torch

from .._C import check, lib
from ..utils import on_device_of, torch_dtype_code


@on_device_of("q")
def attention_fp16(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, scale: float) -> torch.Tensor:
    """Writes ``o`` ([B, T_q, H * 128], fp16 or bf16) in place and returns it."""
    if not q.is_cuda:
        raise RuntimeError("nunchaku_b200 has no CPU path: q must be a CUDA tensor")
    for name, t in (("q", q), ("k", k), ("v", v)):
        if t.dtype != torch.float16 or t.dim() != 4 or t.shape[-1] != 128 or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous fp16 [B, H, T, 128] tensor")
    B, H, Tq, _ = q.shape
    if k.shape != v.shape or k.shape[0] != B or k.shape[1] != H:
        raise ValueError("k / v must be [B, H, T_kv, 128] with q's batch and heads")
    Tkv = k.shape[2]
    if Tq % 128 or Tkv % 128:
        raise ValueError("token counts must be multiples of 128 (pad like the PackQKV epilogue does)")
    if o.dtype not in (torch.float16, torch.bfloat16) or tuple(o.shape) != (B, Tq, H * 128) or not o.is_contiguous():
        raise ValueError("o must be a contiguous fp16 / bf16 [B, T_q, H * 128] tensor")
    check(lib.nb200_attention_fp16(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), torch_dtype_code(o.dtype), B, H, Tq, Tkv, float(scale),
                                   torch.cuda.current_stream().cuda_stream), "attention_fp16")
    return o
