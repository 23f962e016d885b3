"""Depthwise 3x3 convolution of SANA's GLUMBConv on B200 -- the reference's ``dwconv_f16(input, weight, out, bias)`` (src/kernels/dwconv.h:9;
module ``DWCONV``, src/Linear.cpp:541-551) over ``nb200_dwconv3x3``.  NHWC, stride 1, zero padding 1."""
from __future__ import annotations

import torch
from torch import nn

from .._C import check, lib
from ..utils import on_device_of, torch_dtype_code


@on_device_of("input")
def dwconv_f16(input: torch.Tensor, weight: torch.Tensor, out: torch.Tensor | None = None, bias: torch.Tensor | None = None) -> torch.Tensor:
    """``input`` hT [N, H, W, C]; ``weight`` hT [C, 3, 3, 1] (or [C, 3, 3]); ``bias`` hT [C] or None.  Returns ``out`` ([N, H, W, C], allocated
    when None)."""
    if not input.is_cuda:
        raise RuntimeError("nunchaku_b200 has no CPU path: input must be a CUDA tensor")
    if input.dim() != 4 or input.dtype not in (torch.float16, torch.bfloat16) or not input.is_contiguous():
        raise ValueError("input must be a contiguous fp16 / bf16 [N, H, W, C] tensor")
    N, H, W, C = input.shape
    if weight.dtype != input.dtype or weight.numel() != C * 9 or weight.shape[0] != C or not weight.is_contiguous():
        raise ValueError("weight must be a contiguous [C, 3, 3, 1] tensor of the input dtype")
    if bias is not None and (bias.dtype != input.dtype or bias.numel() != C or not bias.is_contiguous()):
        raise ValueError("bias must be a contiguous [C] tensor of the input dtype")
    if C % 8:
        raise ValueError("C must be a multiple of 8")
    if out is None:
        out = torch.empty_like(input)
    elif out.shape != input.shape or out.dtype != input.dtype or not out.is_contiguous() or out.data_ptr() == input.data_ptr():
        raise ValueError("out must be a distinct contiguous tensor of input's shape and dtype")
    check(lib.nb200_dwconv3x3(torch_dtype_code(input.dtype), input.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(), out.data_ptr(),
                              N, H, W, C, torch.cuda.current_stream().cuda_stream), "dwconv3x3")
    return out


class DWCONV(nn.Module):
    """The reference module's parameters (``weight`` [C, 3, 3, 1], ``bias`` [C]; src/Linear.cpp:541-547)."""

    def __init__(self, in_features: int, use_bias: bool = True, torch_dtype: torch.dtype = torch.bfloat16, device=None):
        super().__init__()
        self.in_features = in_features
        self.weight = nn.Parameter(torch.empty(in_features, 3, 3, 1, dtype=torch_dtype, device=device), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(in_features, dtype=torch_dtype, device=device), requires_grad=False) if use_bias else None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return dwconv_f16(x, self.weight, None, self.bias)
