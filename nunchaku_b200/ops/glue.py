"""Elementwise / row-reduce glue between the SVDQuant linears.

Host mirror of the reference's C++ helpers -- same names, argument meaning and in-place / allocating
behaviour as ``nunchaku::kernels::{add, mul_add, mul_add_batch, split_mod, cast}``
(src/kernels/misc_kernels.h:8-25), ``Silu/GELU::forward`` (src/activation.cpp:4-14) and
``LayerNorm/RMSNorm::forward`` (src/layernorm.cpp:14-24).  Every function launches one CUDA kernel of
``csrc/glue.cu`` through the C ABI on torch's current stream; there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from .._C import check, lib

_DTYPE = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}
ACT_SILU, ACT_GELU = 1, 2


def _code(t: torch.Tensor) -> int:
    try:
        return _DTYPE[t.dtype]
    except KeyError:
        raise TypeError(f"glue ops support float16 / bfloat16 / float32, got {t.dtype}") from None


def _cuda(*ts: torch.Tensor | None) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("nunchaku_b200 has no CPU path: tensors must be CUDA tensors")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _activation(kind: int, x: torch.Tensor) -> torch.Tensor:
    _cuda(x)
    x = x.contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.nb200_activation(kind, _code(x), x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "nb200_activation")
    return out


def silu(x: torch.Tensor) -> torch.Tensor:
    """``Silu::forward`` (src/activation.cpp:4-8): ``T(x / (1 + expf(-x)))``."""
    return _activation(ACT_SILU, x)


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """``GELU::forward`` (src/activation.cpp:10-14) = vllm ``gelu_new`` with its mixed 16-bit / fp32 chain."""
    return _activation(ACT_GELU, x)


def layernorm(x: torch.Tensor, weight: torch.Tensor | None = None, bias: torch.Tensor | None = None, eps: float = 1e-5) -> torch.Tensor:
    """``LayerNorm::forward`` (src/layernorm.cpp:14-18): normalise the last dim; optional elementwise affine."""
    _cuda(x, weight, bias)
    x = x.contiguous()
    hidden = x.shape[-1]
    for p in (weight, bias):
        if p is not None and (p.dtype != x.dtype or p.numel() != hidden or not p.is_contiguous()):
            raise ValueError("weight / bias must be contiguous [hidden] tensors of x's dtype")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.nb200_layernorm(_code(x), x.data_ptr(), _ptr(weight), _ptr(bias), out.data_ptr(), x.numel() // hidden if hidden else 0,
                                  hidden, float(eps), _stream()), "nb200_layernorm")
    return out


def layernorm_mod(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, eps: float = 1e-6, *, scale_shift: float = 1.0,
                  weight: torch.Tensor | None = None, bias: torch.Tensor | None = None) -> torch.Tensor:
    """AdaLN in one pass over the activations: ``LayerNorm(x) * (scale + scale_shift) + shift`` with one modulation vector
    (``scale`` / ``shift``: ``hidden`` elements) for all rows -- what the reference's AdaLayerNormZero does as ``layernorm`` followed by
    ``mul_add_batch(norm, scale, True, 1.0, shift, True)`` (src/FluxModel.cpp:41-96); same rounding points, bit-identical result."""
    _cuda(x, scale, shift, weight, bias)
    x = x.contiguous()
    hidden = x.shape[-1]
    for p in (scale, shift, weight, bias):
        if p is not None and (p.dtype != x.dtype or p.numel() != hidden or not p.is_contiguous()):
            raise ValueError("scale / shift / weight / bias must be contiguous tensors of hidden elements and x's dtype")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.nb200_layernorm_mod(_code(x), x.data_ptr(), _ptr(weight), _ptr(bias), scale.data_ptr(), shift.data_ptr(), float(scale_shift),
                                      out.data_ptr(), x.numel() // hidden if hidden else 0, hidden, float(eps), _stream()), "nb200_layernorm_mod")
    return out


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """``RMSNorm::forward`` with ``use_quant = false`` (src/layernorm.cpp:20-24)."""
    _cuda(x, weight)
    x = x.contiguous()
    hidden = x.shape[-1]
    if weight.dtype != x.dtype or weight.numel() != hidden or not weight.is_contiguous():
        raise ValueError("weight must be a contiguous [hidden] tensor of x's dtype")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.nb200_rms_norm(_code(x), x.data_ptr(), weight.data_ptr(), out.data_ptr(), x.numel() // hidden if hidden else 0, hidden,
                                 float(eps), _stream()), "nb200_rms_norm")
    return out


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """``kernels::add`` (misc_kernels.cu:7-27): same shape, same dtype, contiguous."""
    _cuda(a, b)
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("add: shapes and dtypes must match")
    if not (a.is_contiguous() and b.is_contiguous()):
        raise ValueError("add: tensors must be contiguous (misc_kernels.cu:10-11)")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        check(lib.nb200_add(_code(a), a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "nb200_add")
    return out


def mul_add_batch(x: torch.Tensor, scale: torch.Tensor | None, batch_scale: bool, scale_shift: float, bias: torch.Tensor,
                  batch_bias: bool) -> None:
    """``kernels::mul_add_batch`` (misc_kernels.cu:70-131), in place on ``x``.

    ``x[b] = x[b] * (scale[b or 0] + scale_shift) + bias[b or 0]`` with scale / bias broadcast cyclically over
    each batch item's elements; ``scale=None`` means ``x += bias``.
    """
    _cuda(x, scale, bias)
    if x.dim() < 1 or not x[0].is_contiguous():
        raise ValueError("mul_add_batch: x must be contiguous within a batch item")
    batch = x.shape[0]
    if batch_scale and scale is not None and scale.shape[0] != batch:
        raise ValueError("mul_add_batch: scale.shape[0] must equal the batch size")
    if batch_bias and bias.shape[0] != batch:
        raise ValueError("mul_add_batch: bias.shape[0] must equal the batch size")
    if bias.dtype != x.dtype or (scale is not None and scale.dtype != x.dtype):
        raise ValueError("mul_add_batch: dtypes must match")
    numel = x.numel() // batch if batch else 0
    numel_scale = (scale.numel() // (batch if batch_scale else 1)) if scale is not None else 1
    numel_bias = bias.numel() // (batch if batch_bias else 1)
    with torch.cuda.device(x.device):
        check(lib.nb200_mul_add_batch(_code(x), x.data_ptr(), _ptr(scale), bias.data_ptr(), float(scale_shift), batch, numel, numel_scale,
                                      numel_bias, x.stride(0) if batch > 1 else numel,
                                      scale.stride(0) if (scale is not None and batch_scale) else 0,
                                      bias.stride(0) if batch_bias else 0, _stream()), "nb200_mul_add_batch")


def mul_add(x: torch.Tensor, scale: torch.Tensor | None, bias: torch.Tensor) -> None:
    """``kernels::mul_add`` (misc_kernels.cu:29-68): ``x = x * scale + bias`` in place, scale / bias cyclic."""
    _cuda(x, scale, bias)
    if not x.is_contiguous():
        raise ValueError("mul_add: x must be contiguous")
    if bias.dtype != x.dtype or (scale is not None and scale.dtype != x.dtype):
        raise ValueError("mul_add: dtypes must match")
    with torch.cuda.device(x.device):
        check(lib.nb200_mul_add_batch(_code(x), x.data_ptr(), _ptr(scale), bias.data_ptr(), 0.0, 1, x.numel(),
                                      scale.numel() if scale is not None else 1, bias.numel(), 0, 0, 0, _stream()), "nb200_mul_add_batch")


def split_mod(input: torch.Tensor, n: int) -> list[torch.Tensor]:
    """``kernels::split_mod<N>`` (misc_kernels.cu:187-214): de-interleave the last dim into ``n`` tensors."""
    _cuda(input)
    if not 2 <= n <= 6:
        raise ValueError("split_mod: n must be in 2..6")
    if input.shape[-1] % n != 0:
        raise ValueError("split_mod: last dim must be a multiple of n")
    input = input.contiguous()
    shape = list(input.shape)
    shape[-1] //= n
    outs = [torch.empty(shape, dtype=input.dtype, device=input.device) for _ in range(n)]
    arr = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    with torch.cuda.device(input.device):
        check(lib.nb200_split_mod(_code(input), input.data_ptr(), arr, n, input.numel(), _stream()), "nb200_split_mod")
    return outs


def cast(input: torch.Tensor, output: torch.Tensor) -> None:
    """``kernels::cast`` (misc_kernels.cu:256-285): ``output[...] = input[...]`` with dtype conversion."""
    _cuda(input, output)
    if input.shape != output.shape:
        raise ValueError("cast: shapes must match")
    if not (input.is_contiguous() and output.is_contiguous()):
        raise ValueError("cast: tensors must be contiguous")
    with torch.cuda.device(input.device):
        check(lib.nb200_cast(_code(input), input.data_ptr(), _code(output), output.data_ptr(), input.numel(), _stream()), "nb200_cast")
