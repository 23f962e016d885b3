"""AWQ W4A16 GEMV on B200 -- the reference's operator signature (nunchaku/ops/gemv.py:10-58) over ``nb200_gemv_awq``, and the
layer built on it (the AdaLN modulation linears of FLUX: ``AWQW4A16Linear``, nunchaku/models/linear.py:277-420)."""
from __future__ import annotations

import torch
from torch import nn

from .._C import check, lib
from ..utils import on_device_of, torch_dtype_code


@on_device_of("in_feats")
def awq_gemv_w4a16_cuda(in_feats: torch.Tensor, kernel: torch.Tensor, scaling_factors: torch.Tensor, zeros: torch.Tensor, m: int, n: int, k: int,
                        group_size: int = 64, *, bias: torch.Tensor | None = None, fuse_silu: bool = False) -> torch.Tensor:
    """``in_feats`` (k,) or (m, k) hT; ``kernel`` int32 (n // 4, k // 2) in the checkpoint layout (read in place); ``scaling_factors`` /
    ``zeros`` hT (k // group_size, n).  Returns (m, n) hT [(n,) for a 1-D input].  Keyword-only extensions: ``bias`` (n,) is
    added in the kernel's epilogue, ``fuse_silu`` applies SiLU to the input first -- the AdaLN modulation's silu -> gemv -> + bias as ONE launch,
    bit-identical to the three."""
    if not in_feats.is_cuda:
        raise RuntimeError("nunchaku_b200 has no CPU path: in_feats must be a CUDA tensor")
    x = in_feats.contiguous().view(-1, k)
    if x.shape[0] != m:
        raise ValueError("m does not match in_feats")
    if kernel.dtype != torch.int32 or kernel.numel() * 8 != n * k or not kernel.is_contiguous():
        raise ValueError("kernel must be a contiguous int32 tensor of n * k / 8 elements")
    for t in (scaling_factors, zeros):
        if t.dtype != x.dtype or tuple(t.shape) != (k // group_size, n) or not t.is_contiguous():
            raise ValueError("scaling_factors / zeros must be contiguous (k // group_size, n) tensors of the input dtype")
    if bias is not None and (bias.dtype != x.dtype or bias.numel() != n or not bias.is_contiguous()):
        raise ValueError("bias must be a contiguous (n,) tensor of the input dtype")
    out = torch.empty(m, n, dtype=x.dtype, device=x.device)
    check(lib.nb200_gemv_awq_fused(torch_dtype_code(x.dtype), x.data_ptr(), kernel.data_ptr(), scaling_factors.data_ptr(), zeros.data_ptr(),
                                   None if bias is None else bias.data_ptr(), out.data_ptr(), m, n, k, group_size, int(bool(fuse_silu)),
                                   torch.cuda.current_stream().cuda_stream), "gemv_awq")
    return out.view(n) if in_feats.dim() == 1 else out


class AWQW4A16Linear(nn.Module):
    """State-dict compatible with the reference layer (qweight int32 (out // 4, in // 2), wscales / wzeros (in // G, out), bias)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, group_size: int = 64, torch_dtype: torch.dtype = torch.bfloat16,
                 device: str | torch.device | None = None):
        super().__init__()
        device = torch.device("cpu") if device is None else device
        self.in_features, self.out_features, self.group_size = in_features, out_features, group_size
        groups = in_features // group_size
        self.qweight = nn.Parameter(torch.empty(out_features // 4, in_features // 2, dtype=torch.int32, device=device), requires_grad=False)
        self.wscales = nn.Parameter(torch.empty(groups, out_features, dtype=torch_dtype, device=device), requires_grad=False)
        self.wzeros = nn.Parameter(torch.empty(groups, out_features, dtype=torch_dtype, device=device), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(out_features, dtype=torch_dtype, device=device)) if bias else None

    @classmethod
    def from_linear(cls, linear: nn.Linear, group_size: int = 64, torch_dtype: torch.dtype = torch.bfloat16, device="cpu", **_):
        return cls(linear.in_features, linear.out_features, bias=linear.bias is not None, group_size=group_size, torch_dtype=torch_dtype, device=device)

    def forward(self, x: torch.Tensor, *, fuse_silu: bool = False) -> torch.Tensor:
        """``fuse_silu``: y = W silu(x) + b in one launch (what AdaLayerNormZero does with three, src/FluxModel.cpp:41-96)"""
        return awq_gemv_w4a16_cuda(x, self.qweight, self.wscales, self.wzeros, x.shape[0], self.out_features, self.in_features, self.group_size,
                                   bias=self.bias, fuse_silu=fuse_silu)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, group_size={self.group_size}"
