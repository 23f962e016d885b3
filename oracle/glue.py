"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's elementwise / row-reduce glue
(SURVEY.md section 8, row a14).  Only tests/, __graft_entry__.smoke() and bench.py may import this.

PARITY UNPINNED for the same reason as oracle/svdq.py: the reference implements these only as CUDA
kernels with no CPU twin and no golden vectors.  Each function follows the CUDA source op by op,
keeping every rounding point (16-bit ``T`` arithmetic vs fp32):

  silu, gelu_new        src/kernels/activation_kernels_impl.cuh:7-10, 93-97
  layernorm             src/kernels/layernorm_kernels_impl.cuh:13-22, 46-164 (USE_DIFF_OF_SQUARES = true,
                        layernorm_kernels.cu:36-58)
  rms_norm              src/kernels/layernorm_kernels_impl.cuh:292-320
  add, mul_add(_batch), split_mod, cast
                        src/kernels/misc_kernels_impl.cuh:13-90, 183-203; misc_kernels.cu:70-131

Two deliberate modelling choices, both stated in the tests:
  * ``x * (scale + shift) + bias`` in 16-bit types: CUDA's half / bfloat16 ``operator*`` and
    ``operator+`` lower to contractable ``mul`` / ``add`` PTX, which ptxas fuses into one HFMA2 (checked in the
    SASS of our kernel, built from the same expression) -> ONE rounding of the exact x*s+b.
  * row statistics are fp32 sums whose order is implementation-defined (block reduce); the oracle sums in
    float64 and rounds once, so norm outputs are compared with a 1-ulp-of-T tolerance, not bit equality.
"""
from __future__ import annotations

import numpy as np
import torch

from .svdq import rn

_F32 = torch.float32


def _t(x: torch.Tensor, dtype) -> torch.Tensor:
    """float32 -> T (round to nearest even); identity for fp32."""
    return x.to(dtype)


def _clamp_half(x: torch.Tensor) -> torch.Tensor:
    if x.dtype == torch.float16:
        return x.clamp(-65504.0, 65504.0)
    return x


def _t_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a * b in T: the exact product of two <= 11-bit significands fits float32, one rounding to T."""
    if a.dtype == _F32:
        return a * b
    return (a.to(_F32) * b.to(_F32)).to(a.dtype)


def _t_add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    if a.dtype == _F32:
        return a + b
    return rn(a.to(torch.float64) + b.to(torch.float64), a.dtype)


def silu(x: torch.Tensor) -> torch.Tensor:
    f = x.to(_F32)
    return _t(f / (1.0 + torch.exp(-f)), x.dtype)


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    T = x.dtype
    x3 = _t_mul(_t_mul(x, x), x).to(_F32)
    inner = _t_add(x, _t(torch.tensor(0.044715, dtype=_F32) * x3, T))
    arg = _t(torch.tensor(0.79788456, dtype=_F32) * inner.to(_F32), T)
    t = _t(torch.tanh(arg.to(_F32)), T)
    half = torch.tensor(0.5, dtype=T)
    one = torch.tensor(1.0, dtype=T)
    return _t_mul(_t_mul(half.expand_as(x), x), _t_add(one.expand_as(t).contiguous(), t))


def layernorm(x: torch.Tensor, weight: torch.Tensor | None, bias: torch.Tensor | None, eps: float) -> torch.Tensor:
    H = x.shape[-1]
    f = x.to(torch.float64)
    mean = (f.sum(-1, keepdim=True) / H).to(_F32)
    var = (f.square().sum(-1, keepdim=True) / H).to(_F32) - mean * mean
    rstd = (1.0 / torch.sqrt((var + np.float32(eps)).to(torch.float64))).to(_F32)
    r = (x.to(_F32) - mean) * rstd
    if weight is not None:
        r = r * weight.to(_F32)
    if bias is not None:
        r = r + bias.to(_F32)
    return _t(r, x.dtype)


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    H = x.shape[-1]
    ms = (x.to(torch.float64).square().sum(-1, keepdim=True) / H).to(_F32)
    rstd = (1.0 / torch.sqrt((ms + np.float32(eps)).to(torch.float64))).to(_F32)
    n = _t(x.to(_F32) * rstd, x.dtype)
    return _t_mul(n, weight.expand_as(n).contiguous())


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return _t_add(a, b)


def mul_add_batch(x: torch.Tensor, scale: torch.Tensor | None, batch_scale: bool, scale_shift: float, bias: torch.Tensor,
                  batch_bias: bool) -> torch.Tensor:
    """Returns the new x (the reference works in place)."""
    T = x.dtype
    B = x.shape[0]
    xf = x.reshape(B, -1)
    n = xf.shape[1]

    def tile(v: torch.Tensor, batched: bool) -> torch.Tensor:
        v = v.reshape(B, -1) if batched else v.reshape(1, -1).expand(B, -1)
        return v.repeat(1, n // v.shape[1])

    bb = tile(bias, batch_bias)
    if scale is None:
        out = _t_add(xf.contiguous(), bb.contiguous())
    else:
        s = _t_add(tile(scale, batch_scale).contiguous(), torch.tensor(scale_shift, dtype=T).expand(B, n).contiguous())
        out = rn(xf.to(torch.float64) * s.to(torch.float64) + bb.to(torch.float64), T)  # fused multiply-add, one rounding
    return _clamp_half(out).reshape(x.shape)


def split_mod(x: torch.Tensor, n: int) -> list[torch.Tensor]:
    shape = list(x.shape)
    shape[-1] //= n
    flat = x.reshape(-1, n)
    return [flat[:, k].reshape(shape).contiguous() for k in range(n)]


def cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return _clamp_half(x.to(_F32).to(dtype))


# --------------------------------------------------------------------------------------------------
# SANA linear attention (row a13): EpilogueLiteLA second half + vk_mul_q  (epilogues.cuh:552-760)
# --------------------------------------------------------------------------------------------------
def litela_vk(qkv: torch.Tensor):
    """qkv hT [B, T, N], N = 3 * heads * 32 laid out [Q | per head: K(32) V(32)] ->
    (relu(Q) hT [B, T, N/3], vk float32 [B, heads, 33, 32]) with fp64 accumulation (the device sums fp32 in an
    unspecified order)."""
    B, T, N = qkv.shape
    heads = N // 96
    q = torch.clamp_min(qkv[..., : N // 3].to(_F32), 0).to(qkv.dtype)
    kv = qkv[..., N // 3:].to(torch.float64).reshape(B, T, heads, 2, 32)
    k = torch.clamp_min(kv[:, :, :, 0], 0)     # [B, T, H, 32]
    v = kv[:, :, :, 1]
    vk = torch.einsum("bthv,bthk->bhvk", v, k)
    ksum = k.sum(1).unsqueeze(2)                # [B, H, 1, 32]
    return q, torch.cat([vk, ksum], 2).to(_F32)


def vk_mul_q(q: torch.Tensor, vk: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """q hT [B, T, heads*32], vk f32 [B, heads, 33, 32] -> new q.  The device accumulates q[i] * vk[j][i] in fp32 with
    fused multiply-adds in index order and divides with div.approx; here fp64 dot products and exact division."""
    B, T, _ = q.shape
    heads = vk.shape[1]
    qh = q.reshape(B, T, heads, 32).to(torch.float64)
    out = torch.einsum("bthi,bhji->bthj", qh, vk.to(torch.float64))       # [B, T, H, 33]
    res = out[..., :32] / (out[..., 32:33].to(_F32).to(torch.float64) + np.float32(eps))
    return res.to(_F32).to(q.dtype).reshape(B, T, heads * 32)
