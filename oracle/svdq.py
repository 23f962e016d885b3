"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the SVDQuant W4A4 fused-linear arithmetic.

PARITY UNPINNED: the reference implements this path only as CUDA/PTX kernels
(src/kernels/zgemm/*.cuh), ships no CPU implementation, no operator-level test and no
golden vector for it; it can neither be run in the authoring
container (no GPU) nor on the B200 box (its loader rejects sm_100 and the NVFP4 kernel does
not assemble for sm_100a).  Every function below therefore follows the CUDA source line by
line and cites it; formats are pinned separately (oracle/formats.py).

Everything here works on LOGICAL (un-swizzled) tensors:

  x            [M, K]      fp16/bf16 activations ("hT" = the model's 16-bit dtype)
  smooth       [K]         hT
  lora_down    [R, K]      hT      (reference stores it packed as [K, R])
  qa           [Mp, K]     int8    INT4: signed -8..7 or unsigned 0..15 values
                                   NVFP4: e2m1 codes 0..15 (bit3 = sign)
  ascales      [K/G, Mp]   hT (INT4, G=64)  |  uint8 e4m3 bit patterns (NVFP4, G=16)
  lora_act     [Mp, R]     fp32
  qw           [N, K]      int8 (same coding as qa; weights are always signed)
  wscales      [N, K/G]    hT | uint8 e4m3
  bias, wcscales [N]       hT
  lora_up      [N, R]      hT

Two arithmetic modes for the GEMM:
  * ``mode="ref"``   -- emulates the reference kernel's roundings (16-bit accumulator
                        chain for INT4, fp32 block-scaled sum for NVFP4);
  * ``mode="exact"`` -- fp64 end to end, one final rounding.

Approximate PTX primitives used by the reference (rcp.approx, div.approx, tanh.approx,
ex2.approx, rsqrt.approx) are evaluated exactly here; see ``compare_codes`` for the
resulting comparison rule on 4-bit codes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

HALF_DTYPES = (torch.float16, torch.bfloat16)

# e2m1 magnitudes by code (bit 3 is the sign)
E2M1_VALUES = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=torch.float64)
SHIFT_GELU = 0.171875  # gemm_w4a4_launch_impl.cuh:286


def ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


# --------------------------------------------------------------------------------------
# correctly-rounded narrowing from float64
# --------------------------------------------------------------------------------------
def _f64_to_f32_round_to_odd(x: torch.Tensor) -> torch.Tensor:
    """float64 -> float32 with round-to-odd, so that a following RNE to <=22 bits equals a
    single correct rounding of the float64 value."""
    a = x.detach().to(torch.float64).contiguous().numpy()
    with np.errstate(over="ignore", invalid="ignore"):
        f = a.astype(np.float32)
    back = f.astype(np.float64)
    inexact = np.isfinite(a) & np.isfinite(back) & (back != a)
    away = inexact & (np.abs(back) > np.abs(a))
    f = np.where(away, np.nextafter(f, np.float32(0.0)), f).astype(np.float32)
    bits = f.view(np.uint32)
    bits = np.where(inexact, bits | np.uint32(1), bits)
    return torch.from_numpy(bits.view(np.float32).copy())


def rn(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Round a float64 (or float32) tensor to ``dtype`` (fp16/bf16/fp32) with ONE
    round-to-nearest-even, the way a fused hardware op does."""
    if x.dtype != torch.float64:
        return x.to(dtype)
    if dtype == torch.float64:
        return x
    if dtype == torch.float32:
        return x.to(torch.float32)  # IEEE double->float is correctly rounded
    if dtype == torch.float16:
        with np.errstate(over="ignore", invalid="ignore"):
            return torch.from_numpy(x.contiguous().numpy().astype(np.float16))
    if dtype == torch.bfloat16:
        return _f64_to_f32_round_to_odd(x).to(torch.bfloat16)
    raise TypeError(dtype)


def f64(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float64)


# --------------------------------------------------------------------------------------
# 4-bit / 8-bit float codecs
# --------------------------------------------------------------------------------------
def e2m1_decode(codes: torch.Tensor) -> torch.Tensor:
    """codes 0..15 -> float64 value."""
    c = codes.to(torch.int64) & 0xF
    mag = E2M1_VALUES.to(codes.device)[c & 7]
    return torch.where((c & 8) != 0, -mag, mag)


def e2m1_encode(v: torch.Tensor) -> torch.Tensor:
    """``cvt.rn.satfinite.e2m1x2.f32`` (gemm_utils.cuh:239-245): round-to-nearest-even
    onto {0,.5,1,1.5,2,3,4,6}, saturate to +-6, NaN -> +6 (format has no NaN)."""
    v = v.to(torch.float64)
    a = v.abs()
    # thresholds are the midpoints; ties go to the code with even mantissa bit
    #  value:   0   .5   1   1.5   2    3    4    6
    #  code :   0    1   2    3    4    5    6    7
    code = torch.zeros_like(a, dtype=torch.int64)
    code = torch.where(a > 0.25, torch.ones_like(code), code)          # tie .25 -> 0 (even)
    code = torch.where(a >= 0.75, torch.full_like(code, 2), code)      # tie .75 -> 1.0 (code 2, even)
    code = torch.where(a > 1.25, torch.full_like(code, 3), code)       # tie 1.25 -> 1.0
    code = torch.where(a >= 1.75, torch.full_like(code, 4), code)      # tie 1.75 -> 2.0
    code = torch.where(a > 2.5, torch.full_like(code, 5), code)        # tie 2.5 -> 2.0
    code = torch.where(a >= 3.5, torch.full_like(code, 6), code)       # tie 3.5 -> 4.0
    code = torch.where(a > 5.0, torch.full_like(code, 7), code)        # tie 5.0 -> 4.0
    sign = (torch.signbit(v) & ~torch.isnan(v)).to(torch.int64) << 3
    code = code | sign
    code = torch.where(torch.isnan(v), torch.full_like(code, 7), code)
    return code.to(torch.int8)


def e4m3_decode(bits: torch.Tensor) -> torch.Tensor:
    return bits.contiguous().view(torch.uint8).view(torch.float8_e4m3fn).to(torch.float64)


def e4m3_encode(v32: torch.Tensor) -> torch.Tensor:
    """``cvt.rn.satfinite.e4m3x2.f32`` (gemm_utils.cuh:247-252) for non-negative finite
    inputs <= 448 (the caller clamps): returns uint8 bit patterns."""
    v = v32.to(torch.float32).clamp(min=-448.0, max=448.0)
    return v.to(torch.float8_e4m3fn).view(torch.uint8)


# --------------------------------------------------------------------------------------
# elementwise pieces
# --------------------------------------------------------------------------------------
def gelu_tanh_f32(x: torch.Tensor) -> torch.Tensor:
    """gemm_utils.cuh:306-312 (gelu_half2): fp32 math, tanh.approx -> exact tanh here."""
    xf = x.to(torch.float64)
    t = 0.5 + 0.5 * torch.tanh(0.79788456 * (xf + 0.044715 * xf * xf * xf))
    return xf * t


def silu_f32(x: torch.Tensor) -> torch.Tensor:
    """gemm_utils.cuh:291-304,322-327: x * 1/(1+2^(-x*log2e))."""
    xf = x.to(torch.float64)
    return xf / (1.0 + torch.exp(-xf))


def h_div(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """h2div (gemm_utils.cuh:329-344): fp32 __fdividef then round to hT."""
    q = a.to(torch.float32) / b.to(torch.float32)
    return q.to(a.dtype)


# --------------------------------------------------------------------------------------
# activation quantisation
# --------------------------------------------------------------------------------------
def _quantize_rows_int4(xs: torch.Tensor, unsigned: bool):
    """quantize_w4a4_from_fpsum_warp (gemm_w4a4.cuh:429-523).
    xs: [M, K] hT (already smoothed / shifted).  Returns (q int8 [M,K], scales hT [K/64, M])."""
    M, K = xs.shape
    assert K % 64 == 0
    hT = xs.dtype
    g = xs.view(M, K // 64, 64)
    amax = g.abs().amax(dim=-1)                                   # hT max, exact (:467-480)
    recip = torch.tensor(1.0 / 15.0 if unsigned else 1.0 / 7.0, dtype=torch.float32)
    s32 = amax.to(torch.float32) * recip                          # :485-486 fp32 multiply
    scales = s32.to(hT)                                           # :487-490
    with np.errstate(divide="ignore"):
        rs = (1.0 / s32.to(torch.float64)).to(torch.float32)      # rcp.approx.ftz (:495-496)
    prod = g.to(torch.float32) * rs.unsqueeze(-1)                 # :504 fp32 multiply
    q = torch.round(prod)                                         # cvt.rni
    q = torch.nan_to_num(q, nan=0.0)                              # cvt of NaN -> 0
    lo, hi = (0, 15) if unsigned else (-8, 7)                     # cvt.pack.sat.{u4,s4}
    q = q.clamp(lo, hi).to(torch.int8).view(M, K)
    return q, scales.t().contiguous()


def _quantize_rows_fp4(xs: torch.Tensor):
    """quantize_w4a4_fp4_from_fpsum_warp (gemm_w4a4.cuh:85-187).
    Returns (e2m1 codes int8 [M,K], e4m3 bit patterns uint8 [K/16, M])."""
    M, K = xs.shape
    assert K % 16 == 0
    g = xs.view(M, K // 16, 16)
    amax = g.abs().amax(dim=-1)
    s32 = torch.minimum(amax.to(torch.float32) * torch.tensor(1.0 / 6.0, dtype=torch.float32),
                        torch.tensor(448.0))                      # :133-134
    sbits = e4m3_encode(s32)                                      # :141-142
    with np.errstate(divide="ignore"):
        rs = (1.0 / s32.to(torch.float64)).to(torch.float32)      # :137-138 (unrounded scale)
    prod = g.to(torch.float32) * rs.unsqueeze(-1)                 # :161 (0*inf -> NaN -> +6)
    codes = e2m1_encode(prod).view(M, K)
    return codes, sbits.t().contiguous()


@dataclass
class QuantizedAct:
    q: torch.Tensor          # [Mp, K] int8
    scales: torch.Tensor     # [K/G, Mp]
    lora_act: torch.Tensor   # [Mp, R] fp32
    M: int                   # valid rows


def quantize_w4a4_act_fuse_lora(x: torch.Tensor, smooth: torch.Tensor | None, lora_down: torch.Tensor,
                                *, fp4: bool = False, fuse_glu: bool = False, pad_size: int = 256) -> QuantizedAct:
    """quantize_w4a4_fuse_lora_kernel (gemm_w4a4.cuh:1097-1184) on logical tensors.

    1. load_act_to_fpsum<fuse_glu> (gemm_base.cuh:592-646): rows >= M are zero; GLU is
       ``x[:,2j] * silu(x[:,2j+1])`` with both the silu and the product rounded to hT.
    2. EpilogueLoraDown (lora.cuh:243-353): lora_act = x' @ lora_down^T on the UN-smoothed
       tile, fp32 accumulation (order unspecified: atomics).
    3. EpilogueQuantize<false,false,fp4> with shift 0: xs = hT(x'/smooth), then per group.
    """
    assert x.dtype in HALF_DTYPES and x.dim() == 2
    hT = x.dtype
    M = x.shape[0]
    if fuse_glu:
        a, b = x[:, 0::2], x[:, 1::2]
        sb = rn(silu_f32(b), hT)                                   # silu() returns hT
        x = rn(f64(a) * f64(sb), hT)                               # hT * hT
    K = x.shape[1]
    Mp = ceil_div(M, pad_size) * pad_size
    xp = torch.zeros(Mp, K, dtype=hT)
    xp[:M] = x
    lora_act = (f64(xp) @ f64(lora_down).t()).to(torch.float32)
    xs = h_div(xp, smooth.view(1, K)) if smooth is not None else xp
    if fp4:
        q, s = _quantize_rows_fp4(xs)
    else:
        q, s = _quantize_rows_int4(xs, unsigned=False)
    return QuantizedAct(q=q, scales=s, lora_act=lora_act, M=M)


# --------------------------------------------------------------------------------------
# dequantisation helpers (fp64)
# --------------------------------------------------------------------------------------
def dequant(q: torch.Tensor, scales_rows_by_group: torch.Tensor, fp4: bool) -> torch.Tensor:
    """q [R, K] codes, scales [R, K/G] (hT or e4m3 bits) -> float64 [R, K]."""
    Rr, K = q.shape
    if fp4:
        v = e2m1_decode(q).view(Rr, K // 16, 16)
        s = e4m3_decode(scales_rows_by_group).view(Rr, K // 16, 1)
    else:
        v = q.to(torch.float64).view(Rr, K // 64, 64)
        s = f64(scales_rows_by_group).view(Rr, K // 64, 1)
    return (v * s).view(Rr, K)


# --------------------------------------------------------------------------------------
# the fused GEMM
# --------------------------------------------------------------------------------------
@dataclass
class GemmResult:
    out: torch.Tensor | None = None            # [M, N] hT
    qout: torch.Tensor | None = None           # [Mp, N] int8 codes for the next layer
    oscales: torch.Tensor | None = None        # [N/G, Mp]
    lora_act_out: torch.Tensor | None = None   # [Mp, R2] fp32
    pre_act: torch.Tensor | None = None        # value entering the mid activation (debug)


def _main_int4(qa, ascales, qw, wscales, hT, mode):
    """gemm_w4a4_block (gemm_w4a4.cuh:831-928) + apply_scales (gemm_base.cuh:368-409).
    ref : acc_hT = fma_hT(hT(int32 p), mul_hT(as, ws), acc_hT) per 64-wide group
    exact: sum_g p * as * ws in fp64."""
    Mp, K = qa.shape
    N = qw.shape[0]
    G = K // 64
    a_g = qa.to(torch.float32).view(Mp, G, 64)
    w_g = qw.to(torch.float32).view(N, G, 64)
    if mode == "exact":
        acc = torch.zeros(Mp, N, dtype=torch.float64)
        for g in range(G):
            p = (a_g[:, g] @ w_g[:, g].t()).to(torch.float64)             # exact small ints
            acc += p * f64(ascales[g]).view(Mp, 1) * f64(wscales[:, g]).view(1, N)
        return acc
    acc = torch.zeros(Mp, N, dtype=hT)
    for g in range(G):
        p = a_g[:, g] @ w_g[:, g].t()                                     # |p| <= 64*8*15 < 2^24
        p_h = p.to(hT)                                                    # int2half2: cvt f32 then hT
        sc = rn(f64(ascales[g]).view(Mp, 1) * f64(wscales[:, g]).view(1, N), hT)   # __hmul2
        acc = rn(f64(p_h) * f64(sc) + f64(acc), hT)                       # __hfma2
    return acc


def _main_fp4(qa, ascales, qw, wscales, alpha, hT, mode):
    """gemm_w4a4_fp4_block (gemm_w4a4.cuh:273-356): hardware block-scaled MMA, fp32
    accumulate (internal order unspecified -> fp64 here), then *alpha, then hT."""
    a = dequant(qa, ascales.t(), fp4=True)
    w = dequant(qw, wscales, fp4=True)
    acc = a @ w.t()
    if mode == "exact":
        return acc * float(alpha)
    acc32 = acc.to(torch.float32)
    if float(alpha) != 1.0:
        acc32 = acc32 * torch.tensor(float(alpha), dtype=torch.float32)   # :341-349
    return acc32.to(hT)                                                   # packed_fp32_to_fp16 :351


def rmsnorm_rope(y: torch.Tensor, norm_q: torch.Tensor, norm_k: torch.Tensor, rope_sin: torch.Tensor,
                 rope_cos: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """EpilogueRMSNormRope (epilogues.cuh:269-425) on logical tensors.
    y [M, N] float64 holding hT values; first third of N = Q heads, second third = K heads,
    last third (V) untouched.  rope_sin/cos: [M, 64] fp32 per (row, pair i) acting on columns
    (2i, 2i+1) of each 128-wide head.  fp32 math in the reference; fp64 here."""
    M, N = y.shape
    assert N % 3 == 0 and (N // 3) % 128 == 0
    out = y.clone()
    H = N // 3 // 128
    for part, w in ((0, norm_q), (1, norm_k)):
        blk = y[:, part * (N // 3):(part + 1) * (N // 3)].view(M, H, 128)
        coef = torch.rsqrt((blk * blk).sum(-1, keepdim=True) / 128.0 + eps)   # :327-341
        v = blk * coef * f64(w).view(1, 1, 128)                               # :352-360
        x0, x1 = v[..., 0::2], v[..., 1::2]
        s, c = f64(rope_sin).view(M, 1, 64), f64(rope_cos).view(M, 1, 64)
        r0 = x0 * c - x1 * s                                                  # :362-367
        r1 = x0 * s + x1 * c
        v = torch.stack([r0, r1], dim=-1).view(M, H, 128)
        out[:, part * (N // 3):(part + 1) * (N // 3)] = v.view(M, -1)
    return out


def gemm_w4a4(*, qa: torch.Tensor, ascales: torch.Tensor, qw: torch.Tensor, wscales: torch.Tensor,
              hT: torch.dtype, M: int | None = None, bias: torch.Tensor | None = None,
              lora_act: torch.Tensor | None = None, lora_up: torch.Tensor | None = None,
              lora_scales: list[float] | None = None, fp4: bool = False, alpha: float = 1.0,
              wcscales: torch.Tensor | None = None, act: str = "none",
              next_smooth: torch.Tensor | None = None, next_lora_down: torch.Tensor | None = None,
              want_out: bool = True, rope: tuple | None = None, mode: str = "ref") -> GemmResult:
    """kernels::gemm_w4a4 (zgemm.h:8-36; gemm_w4a4_launch_impl.cuh:7-424) on logical tensors.

    Epilogue order (launch_impl:172-193,196-280,282-423):
      main -> [*alpha] -> wcscale/bias -> LoRA-up -> {none|gelu|silu} -> [LoRA-down(next)]
           -> {store | +0.171875, /smooth, u4 quantise | RMSNorm+RoPE, store}
    ``qa`` holds unsigned values 0..15 when the producer was the fused GELU epilogue
    (act_unsigned) -- the integer dot product is the same code path.
    """
    assert mode in ("ref", "exact")
    Mp, K = qa.shape
    N = qw.shape[0]
    M = Mp if M is None else M
    r = (lambda v: rn(v, hT)) if mode == "ref" else (lambda v: v)

    if fp4:
        acc = _main_fp4(qa, ascales, qw, wscales, alpha, hT, mode)
    else:
        assert float(alpha) == 1.0                                        # launch_impl:107
        acc = _main_int4(qa, ascales, qw, wscales, hT, mode)
    acc = f64(acc)

    # EpilogueBias<USE_BIAS, USE_SCALE> (gemm_base.cuh:710-781): hfma2 / hmul2 / hadd2
    if wcscales is not None and bias is not None:
        acc = f64(r(acc * f64(wcscales).view(1, N) + f64(bias).view(1, N)))
    elif wcscales is not None:
        acc = f64(r(acc * f64(wcscales).view(1, N)))
    elif bias is not None:
        acc = f64(r(acc + f64(bias).view(1, N)))

    # EpilogueLoraUp (lora.cuh:110-241): fp32 accumulate of hT(lora_act*scale) x lora_up
    if lora_up is not None and lora_up.shape[1] > 0:
        R = lora_up.shape[1]
        ls = torch.ones(R // 16, dtype=torch.float32) if lora_scales is None else torch.tensor(
            [float(v) for v in lora_scales][:R // 16] + [0.0] * max(0, R // 16 - len(lora_scales)),
            dtype=torch.float32)                                          # launch_impl:220-225
        la = lora_act.to(torch.float32) * ls.repeat_interleave(16).view(1, R)   # fp32 multiply :149
        la = la.to(hT) if mode == "ref" else la                           # packed_fp32_to_fp16 :151
        add = f64(la) @ f64(lora_up).t()
        acc = acc + add
        if mode == "ref":
            acc = f64(rn(acc.to(torch.float32), hT))                      # fp32 psum -> hT (:221)
    pre_act = acc.clone()

    if act == "gelu":
        acc = f64(r(gelu_tanh_f32(acc)))
    elif act == "silu":
        acc = f64(r(silu_f32(acc)))
    else:
        assert act == "none"

    res = GemmResult(pre_act=pre_act)

    if next_lora_down is not None:                                        # EpilogueLoraDown, no shift
        res.lora_act_out = (acc @ f64(next_lora_down).t()).to(torch.float32)

    if rope is not None:
        norm_q, norm_k, rsin, rcos = rope
        acc = f64(r(rmsnorm_rope(acc, norm_q, norm_k, rsin, rcos)))       # packed_fp32_to_fp16

    if next_smooth is not None:
        # EpilogueQuantize<false, !fp4, fp4> (gemm_w4a4.cuh:930-1043), launch_impl:282-310
        shift = 0.0 if fp4 else SHIFT_GELU
        y = rn(acc + shift, hT)                                           # dst += half2(shift) :981
        ys = h_div(y, next_smooth.to(hT).view(1, N))
        if fp4:
            res.qout, res.oscales = _quantize_rows_fp4(ys)
        else:
            res.qout, res.oscales = _quantize_rows_int4(ys, unsigned=True)

    if want_out:
        out = rn(acc, hT)
        if hT == torch.float16:                                           # gemm_base.cuh:688-696
            out = out.clamp(-65504.0, 65504.0)
        res.out = out[:M]
    return res


# --------------------------------------------------------------------------------------
# comparison rules and synthetic layers
# --------------------------------------------------------------------------------------
def rel_fro(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = f64(a), f64(b)
    d = torch.linalg.norm(b)
    return float(torch.linalg.norm(a - b) / d) if d > 0 else float(torch.linalg.norm(a - b))


def compare_codes(q_test: torch.Tensor, q_ref: torch.Tensor, fp4: bool = False) -> dict:
    """4-bit codes may differ by one step on elements sitting on a rounding boundary,
    because the reference (and our kernel) use rcp.approx / div.approx (<= 2 ulp) where the
    oracle divides exactly (SURVEY.md A.5 tolerance note).  Report mismatch fraction and
    the largest step."""
    if fp4:
        a, b = e2m1_decode(q_test), e2m1_decode(q_ref)
        # step in units of code index on the magnitude grid
        ia = (q_test.to(torch.int64) & 7) * torch.where((q_test.to(torch.int64) & 8) != 0, -1, 1)
        ib = (q_ref.to(torch.int64) & 7) * torch.where((q_ref.to(torch.int64) & 8) != 0, -1, 1)
        diff = (ia - ib).abs()
        mism = (a != b)
    else:
        diff = (q_test.to(torch.int64) - q_ref.to(torch.int64)).abs()
        mism = diff != 0
    return {"frac": float(mism.to(torch.float64).mean()), "max_step": int(diff.max()) if diff.numel() else 0}


@dataclass
class SynthLayer:
    """An SVDQuant-realistic synthetic layer in LOGICAL form."""
    qw: torch.Tensor
    wscales: torch.Tensor
    bias: torch.Tensor
    smooth: torch.Tensor
    lora_down: torch.Tensor   # [R, K]
    lora_up: torch.Tensor     # [N, R]
    wcscales: torch.Tensor | None
    alpha: float
    fp4: bool
    hT: torch.dtype


def make_synthetic_layer(N: int, K: int, R: int, *, fp4: bool = False, hT=torch.bfloat16, seed: int = 0,
                         with_wcscales: bool | None = None) -> SynthLayer:
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(N, K, generator=g, dtype=torch.float64) * 0.02
    n_out = max(1, K // 1000)
    cols = torch.randperm(K, generator=g)[:n_out]
    W[:, cols] *= 20.0
    smooth = torch.exp(torch.randn(K, generator=g, dtype=torch.float64) * 0.5).clamp(0.1, 10.0)
    What = W * smooth.view(1, K)
    if R > 0:
        # randomised range finder instead of a full SVD (fast for 3072^2 on 8 cores)
        Q, _ = torch.linalg.qr(What @ torch.randn(K, R + 8, generator=g, dtype=torch.float64))
        B = Q.t() @ What
        U, S, Vh = torch.linalg.svd(B, full_matrices=False)
        U = (Q @ U)[:, :R]
        lu = (U * S[:R].sqrt().view(1, R)).to(hT)
        # the kernel applies lora_down to the UN-smoothed x (gemm_w4a4.cuh:1159-1171), so the
        # stored factor carries 1/smooth:  (x/smooth) @ (sqrtS Vh)^T == x @ (sqrtS Vh / smooth)^T
        ld = (S[:R].sqrt().view(R, 1) * Vh[:R] / smooth.view(1, K)).to(hT)
        resid = What - f64(lu) @ (f64(ld) * smooth.view(1, K))
    else:
        lu = torch.zeros(N, 0, dtype=hT)
        ld = torch.zeros(0, K, dtype=hT)
        resid = What
    alpha = 1.0
    wcs = None
    if fp4:
        alpha = float(resid.abs().max() / (6.0 * 448.0)) * 4.0 + 1e-8
        rg = (resid / alpha).view(N, K // 16, 16)
        s = (rg.abs().amax(-1) / 6.0).clamp(max=448.0).to(torch.float32)
        sb = e4m3_encode(s)
        sd = e4m3_decode(sb).clamp(min=2.0 ** -9)
        qw = e2m1_encode(rg / sd.unsqueeze(-1)).view(N, K)
        wscales = sb
        if with_wcscales or with_wcscales is None:
            wcs = (1.0 + 0.1 * torch.randn(N, generator=g, dtype=torch.float64)).to(hT)
    else:
        rg = resid.view(N, K // 64, 64)
        s = (rg.abs().amax(-1) / 7.0).to(hT)
        sd = f64(s).clamp(min=1e-30)
        qw = torch.round(rg / sd.unsqueeze(-1)).clamp(-8, 7).to(torch.int8).view(N, K)
        wscales = s
    bias = (torch.randn(N, generator=g, dtype=torch.float64) * 0.1).to(hT)
    return SynthLayer(qw=qw, wscales=wscales, bias=bias, smooth=smooth.to(hT), lora_down=ld, lora_up=lu,
                      wcscales=wcs, alpha=alpha, fp4=fp4, hT=hT)


def make_activations(M: int, K: int, hT=torch.bfloat16, seed: int = 1, smooth: torch.Tensor | None = None,
                     outlier: float = 30.0) -> torch.Tensor:
    """x ~ N(0,1) with 0.1 % outlier channels (x``outlier``).  When ``smooth`` is given the
    outliers are what the smoothing factor absorbs (x = z * smooth, the SmoothQuant
    situation), otherwise they stay in the tensor handed to the quantiser (stress case)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g, dtype=torch.float32)
    n_out = max(1, K // 1000)
    cols = torch.randperm(K, generator=g)[:n_out]
    if smooth is not None:
        x = x * smooth.to(torch.float32).view(1, K)
        x[:, cols] *= 3.0
    else:
        x[:, cols] *= outlier
    return x.to(hT)


def svdq_linear_forward(layer: SynthLayer, x: torch.Tensor, mode: str = "ref", act: str = "none") -> torch.Tensor:
    """SVDQW4A4Linear.forward (nunchaku/models/linear.py:161-188): quantize then GEMM."""
    qa = quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down, fp4=layer.fp4)
    res = gemm_w4a4(qa=qa.q, ascales=qa.scales, qw=layer.qw, wscales=layer.wscales, hT=layer.hT, M=qa.M,
                    bias=layer.bias, lora_act=qa.lora_act, lora_up=layer.lora_up, fp4=layer.fp4,
                    alpha=layer.alpha, wcscales=layer.wcscales, act=act, mode=mode)
    return res.out
