"""Fused SVDQuant W4A4 GEMM on B200.

Two entry points over the same C-ABI call (``nb200_gemm_w4a4``):

* ``svdq_gemm_w4a4_cuda`` -- the reference's operator signature (nunchaku/ops/gemm.py:12-160, 29 positional parameters):
  weight-side tensors arrive in the checkpoint layout and are converted through the ``repack`` cache.  This is what gets
  bound onto the reference's own classes.
* ``gemm_b200`` -- takes a ``B200Weights`` bundle (already converted, owned by ``SVDQW4A4Linear``): no cache lookups on
  the hot path.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field

import torch

from .. import repack
from .._C import NB200_ACT_GELU, NB200_ACT_NONE, NB200_ACT_SILU, NB200_MAX_LORA_SCALES, GemmArgs, check, lib
from ..utils import on_device_of, torch_dtype_code

# tuning knobs for experiments (0 = let the launcher choose)
BLOCK_N_OVERRIDE = 0
NUM_SMS_OVERRIDE = 0
PROF_BUFFER = None   # torch int64 tensor [grid, 16]: per-CTA barrier-wait cycle counters (tools/gemm_prof.py)
# SANA linear attention: True = EpilogueLiteLA inside the GEMM (csrc/gemm_w4a4.cu, EPI_LITELA); False = plain GEMM + nb200_litela_vk; None = by
# measurement (tools/litela_bench.py, profiles/r02i_litela_bench.json, SANA-1.6B's 2 x 1024 x 2304 -> 6912 projection): NVFP4 fused (42 vs 73 us),
# INT4 split (146 vs 155 us: one epilogue warpgroup next to 24 converter warps)
LITELA_FUSED = {"fused": True, "split": False}.get(os.environ.get("NB200_LITELA", "auto"))


@dataclass
class _Call:
    """everything nb200_gemm_args needs, weight side already in B200 layouts"""
    act: torch.Tensor
    ascales: torch.Tensor
    N: int
    K: int
    fp4: bool
    dtype: torch.dtype
    wgt: torch.Tensor = None
    wscales: torch.Tensor = None
    bias: torch.Tensor | None = None
    cscale: torch.Tensor | None = None
    lora_act_in: torch.Tensor | None = None
    lora_up: torch.Tensor | None = None
    rank: int = 0
    lora_scales: list | None = None
    out: torch.Tensor | None = None
    qout: torch.Tensor | None = None
    oscales: torch.Tensor | None = None
    smooth_next: torch.Tensor | None = None
    lora_down_next: torch.Tensor | None = None
    lora_act_out: torch.Tensor | None = None
    rank_down: int = 0
    norm_q: torch.Tensor | None = None
    norm_k: torch.Tensor | None = None
    rotary_emb: torch.Tensor | None = None
    out_qkv: tuple | None = None
    attn_tokens: int = 0
    out_vk: torch.Tensor | None = None      # LiteLA epilogue: `out` is relu(Q) [Mp, N/3]
    vk_tokens: int = 0
    act_unsigned: bool = False
    mid_act: int = NB200_ACT_NONE
    extra: dict = field(default_factory=dict)


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise ValueError(msg)


def _launch(c: _Call) -> None:
    act = c.act
    _require(act.is_contiguous() and act.dtype in (torch.uint8, torch.int8), "act must be a contiguous uint8 tensor [Mp, K/2]")
    Mp = act.numel() // act.shape[-1]
    _require(act.shape[-1] * 2 == c.K, "act and wgt disagree on K")
    _require(c.ascales.is_contiguous(), "ascales must be contiguous")
    if c.fp4:
        _require(c.ascales.element_size() == 1 and c.ascales.numel() == Mp * c.K // 16, "NVFP4 ascales: fp8 [K/16, Mp]")
    else:
        _require(c.ascales.dtype == c.dtype and c.ascales.numel() == Mp * c.K // 64, "INT4 ascales: hT [K/64, Mp]")
    args = GemmArgs()
    args.act, args.ascales = act.data_ptr(), c.ascales.data_ptr()
    args.wgt, args.wscales = c.wgt.data_ptr(), c.wscales.data_ptr()
    args.bias = None if c.bias is None else c.bias.data_ptr()
    args.cscale = None if c.cscale is None else c.cscale.data_ptr()
    rank = c.rank
    if rank > 0:
        la = c.lora_act_in
        _require(la is not None and la.dtype == torch.float32 and la.is_contiguous() and la.shape[-1] == rank and la.numel() == Mp * rank,
                 "lora_act_in must be contiguous float32 [Mp, rank]")
        args.lora_up, args.lora_act_in = c.lora_up.data_ptr(), la.data_ptr()
    scales = c.lora_scales if c.lora_scales is not None else [1.0] * math.ceil(rank / 16)
    for i in range(NB200_MAX_LORA_SCALES):   # absent groups get 0 (launch_impl:220-225)
        args.lora_scales[i] = float(scales[i]) if i < len(scales) else 0.0
    args.Mp, args.N, args.K = Mp, c.N, c.K
    if c.out is not None:
        out2d = c.out.view(-1, c.out.shape[-1])
        _require(out2d.is_contiguous() and out2d.dtype == c.dtype, "out must be a contiguous hT tensor")
        args.out = out2d.data_ptr()
        args.M_out, args.N_out = out2d.shape[0], out2d.shape[1]
    args.R_up, args.R_down = rank, 0
    if c.qout is not None:
        _require(c.qout.is_contiguous() and c.qout.shape[-1] * 2 == c.N and c.qout.numel() // c.qout.shape[-1] == Mp, "qout must be contiguous [Mp, N/2]")
        _require(c.oscales.is_contiguous() and c.oscales.numel() * (16 if c.fp4 else 64) == Mp * c.N, "oscales must hold N/G x Mp scales")
        args.qout, args.oscales = c.qout.data_ptr(), c.oscales.data_ptr()
        args.smooth_next = c.smooth_next.data_ptr()
        if c.rank_down > 0:
            lo = c.lora_act_out
            _require(lo.dtype == torch.float32 and lo.is_contiguous() and tuple(lo.shape) == (Mp, c.rank_down), "lora_act_out must be float32 [Mp, rank]")
            args.lora_down_next, args.lora_act_out, args.R_down = c.lora_down_next.data_ptr(), lo.data_ptr(), c.rank_down
            ws = _reduce_workspace(Mp, c.rank_down, c.act.device)      # deterministic reduction of the per-CTA partial projections
            args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
    keep = []
    if c.rotary_emb is not None:
        rot = c.rotary_emb
        _require(rot.dtype == torch.float32 and rot.is_contiguous() and rot.numel() == Mp * 128, "rotary_emb: packed fp32 [Mp, 128]")
        _require(c.norm_q.numel() == 128 and c.norm_k.numel() == 128 and c.norm_q.dtype == c.dtype and c.norm_k.dtype == c.dtype,
                 "norm_q / norm_k: hT [128]")
        nq, nk = c.norm_q.contiguous(), c.norm_k.contiguous()
        keep += [nq, nk]
        args.rotary_emb, args.norm_q, args.norm_k = rot.data_ptr(), nq.data_ptr(), nk.data_ptr()
        if c.out_qkv is not None:
            heads = c.N // 384
            for t in c.out_qkv:
                _require(t.shape[1] == heads and t.shape[2] >= Mp, f"out_q/out_k/out_v must hold {heads} heads of >= {Mp} rows")
            oq, ok_, ov = c.out_qkv
            args.out_q, args.out_k, args.out_v = oq.data_ptr(), ok_.data_ptr(), ov.data_ptr()
            args.stride_head_q, args.stride_head_k, args.stride_head_v = oq.stride(1), ok_.stride(1), ov.stride(1)
            args.attn_tokens = int(c.attn_tokens)
    if c.out_vk is not None:
        args.out_vk, args.vk_tokens = c.out_vk.data_ptr(), int(c.vk_tokens)
    args.dtype = torch_dtype_code(c.dtype)
    args.fp4 = int(c.fp4)
    args.act_unsigned = int(c.act_unsigned)
    args.mid_act = c.mid_act
    args.block_n = BLOCK_N_OVERRIDE
    args.num_sms = NUM_SMS_OVERRIDE
    args.prof = None if PROF_BUFFER is None else PROF_BUFFER.data_ptr()
    check(lib.nb200_gemm_w4a4(ctypes.byref(args), torch.cuda.current_stream().cuda_stream), "gemm_w4a4")


def _mid_act(fuse_silu: bool, fuse_gelu: bool) -> int:
    if fuse_silu and fuse_gelu:
        raise ValueError("fuse_silu and fuse_gelu are exclusive")
    return NB200_ACT_SILU if fuse_silu else (NB200_ACT_GELU if fuse_gelu else NB200_ACT_NONE)


@on_device_of("act")
def gemm_b200(act: torch.Tensor, ascales: torch.Tensor, lora_act_in: torch.Tensor | None, w, *, out=None, act_unsigned=False,
              lora_scales=None, fuse_silu=False, fuse_gelu=False, next_w=None, qout=None, oscales=None, lora_act_out=None,
              norm_q=None, norm_k=None, rotary_emb=None, out_qkv=None, attn_tokens=0) -> None:
    """The fused GEMM on a converted layer ``w`` (``nunchaku_b200.weights.B200Weights``).

    Mode by outputs, as in the reference (gemm_w4a4_launch_impl.cuh:282,347,407): ``out`` (plain / SiLU / GELU), ``qout`` +
    ``oscales`` (+ ``lora_act_out``) with ``next_w`` = the NEXT layer (fc1 -> GELU -> quantise for fc2), ``rotary_emb`` with
    ``out`` or ``out_qkv`` (RMSNorm + RoPE [+ PackQKV]).
    """
    if not act.is_cuda:
        raise RuntimeError("nunchaku_b200 has no CPU path: act must be a CUDA tensor")
    c = _Call(act=act, ascales=ascales, N=w.N, K=w.K, fp4=w.fp4, dtype=w.dtype, wgt=w.qweight, wscales=w.wscales, bias=w.bias, cscale=w.cscale,
              lora_act_in=lora_act_in, lora_up=w.lora_up, rank=w.rank, lora_scales=lora_scales, out=out, act_unsigned=act_unsigned,
              mid_act=_mid_act(fuse_silu, fuse_gelu), norm_q=norm_q, norm_k=norm_k, rotary_emb=rotary_emb, out_qkv=out_qkv, attn_tokens=attn_tokens)
    if qout is not None:
        _require(next_w is not None and next_w.K == w.N, "the fused quantise epilogue needs the next layer (in_features == this layer's out_features)")
        c.qout, c.oscales, c.smooth_next = qout, oscales, next_w.smooth
        if next_w.rank > 0:
            c.lora_down_next, c.lora_act_out, c.rank_down = next_w.lora_down_next, lora_act_out, next_w.rank
    _launch(c)


@on_device_of("act")
def svdq_gemm_w4a4_cuda(
    act: torch.Tensor,
    wgt: torch.Tensor,
    out: torch.Tensor | None = None,
    qout: torch.Tensor | None = None,
    ascales: torch.Tensor | None = None,
    wscales: torch.Tensor | None = None,
    oscales: torch.Tensor | None = None,
    poolout: torch.Tensor | None = None,
    lora_act_in: torch.Tensor | None = None,
    lora_up: torch.Tensor | None = None,
    lora_down: torch.Tensor | None = None,
    lora_act_out: torch.Tensor | None = None,
    norm_q: torch.Tensor | None = None,
    norm_k: torch.Tensor | None = None,
    rotary_emb: torch.Tensor | None = None,
    bias: torch.Tensor | None = None,
    smooth_factor: torch.Tensor | None = None,
    out_vk: torch.Tensor | None = None,
    out_linearattn: torch.Tensor | None = None,
    act_unsigned: bool = False,
    lora_scales: list[float] | None = None,
    fuse_silu: bool = False,
    fp4: bool = False,
    alpha: float | None = 1.0,
    wcscales: torch.Tensor | None = None,
    out_q: torch.Tensor | None = None,
    out_k: torch.Tensor | None = None,
    out_v: torch.Tensor | None = None,
    attn_tokens: int = 0,
    *,
    fuse_gelu: bool = False,
    qkv_scratch: torch.Tensor | None = None,
):
    """Positional/keyword compatible with the reference wrapper (nunchaku/ops/gemm.py:12-160); results are written in place.
    ``fuse_gelu`` (keyword-only extension) applies the reference's tanh-GELU in the plain epilogue.

    ``act`` / ``ascales`` / ``lora_act_in`` must come from this package's quantize op (or fused epilogue); ``wgt`` /
    ``wscales`` / ``lora_up`` / ``bias`` / ``wcscales`` / ``smooth_factor`` / ``lora_down`` are the reference's packed checkpoint
    tensors, converted once through ``nunchaku_b200.repack`` (see its docstring for invalidation).  ``poolout`` is accepted and
    ignored exactly as in the reference.
    """
    if act is None or wgt is None or ascales is None or wscales is None:
        raise ValueError("act, wgt, ascales and wscales are required")
    if not act.is_cuda:
        raise RuntimeError("nunchaku_b200 has no CPU path: act must be a CUDA tensor")
    litela = out_vk is not None or out_linearattn is not None
    litela_fused = False
    if litela:
        # SANA linear attention (launch_impl:311-346): relu(Q) -> out_linearattn, per-head V^T relu(K) state -> out_vk.
        # Fused route (LITELA_FUSED, the reference's shapes: N / 3 a multiple of 128): the GEMM's own epilogue reduces the K | V tiles on the
        # tensor cores (csrc/gemm_w4a4.cu, EPI_LITELA) and stores relu(Q).  Other shapes: the GEMM writes the plain hT projection into a scratch
        # tensor and nb200_litela_vk reduces it while it is L2-hot.
        _require(out_vk is not None and out_linearattn is not None, "out_vk and out_linearattn go together (launch_impl:313)")
        _require(out_vk.dtype == torch.float32 and out_vk.dim() == 4 and tuple(out_vk.shape[2:]) == (33, 32), "out_vk must be float32 [B, heads, 33, 32]")
        _require(out_linearattn.dim() == 3 and out_linearattn.shape[0] == out_vk.shape[0], "out_linearattn must be [B, tokens, N / 3]")
        _require(out_vk.is_contiguous() and out_linearattn.is_contiguous(), "out_vk / out_linearattn must be contiguous")
        _require(out_linearattn.shape[1] % 256 == 0, "tokens must be a multiple of 256 (launch_impl:331)")
        _require(out_vk.shape[1] * 96 == wgt.shape[0] and out_linearattn.shape[2] * 3 == wgt.shape[0], "N must be 3 * heads * 32")
        litela_fused = (fp4 if LITELA_FUSED is None else LITELA_FUSED) and (wgt.shape[0] // 3) % 128 == 0
        if litela_fused:
            out = out_linearattn.view(-1, out_linearattn.shape[2])
        else:
            out = torch.empty(out_linearattn.shape[0] * out_linearattn.shape[1], wgt.shape[0], dtype=out_linearattn.dtype, device=act.device)
    pack_qkv = out_q is not None or out_k is not None or out_v is not None
    if pack_qkv:
        # EpiloguePackQKV (launch_impl:376-393): fp16 [B=1, heads, rows >= Mp, 128], plain row-major inside a head
        _require(out_q is not None and out_k is not None and out_v is not None, "out_q, out_k and out_v go together")
        _require(rotary_emb is not None, "out_q/out_k/out_v need rotary_emb, norm_q and norm_k (launch_impl:347-376)")
        for t in (out_q, out_k, out_v):
            _require(t.dtype == torch.float16 and t.dim() == 4 and t.shape[0] == 1 and t.shape[-1] == 128, "out_q/out_k/out_v must be float16 [1, heads, tokens_pad, 128]")
            _require(t.stride(-1) == 1 and t.stride(-2) == 128, "out_q/out_k/out_v: the last two dims must be contiguous")
    fused_quant = qout is not None and oscales is not None          # launch_impl:282
    _require((qout is None) == (oscales is None), "qout and oscales go together")
    _require(out is not None or fused_quant or pack_qkv, "out is required unless qout/oscales or out_q/out_k/out_v are given")
    if pack_qkv:
        # the reference ignores `out` in this mode (launch_impl:376-393); `qkv_scratch` (keyword-only extension: hT [Mp, N], contents unspecified
        # afterwards) lets the NVFP4 cluster route run the plain GEMM + the RMSNorm / RoPE / pack kernel instead of the fused epilogue
        out = None
        if qkv_scratch is not None:
            _require(qkv_scratch.dim() == 2 and qkv_scratch.shape[0] == act.numel() // act.shape[-1] and qkv_scratch.shape[1] == wgt.shape[0],
                     "qkv_scratch must be [Mp, N]")
            out = qkv_scratch
    _require(not fused_quant or smooth_factor is not None, "qout needs smooth_factor (the next layer's smoothing vector)")
    _require((lora_down is None) == (lora_act_out is None), "lora_down and lora_act_out go together (launch_impl:199)")
    if lora_down is not None and not fused_quant:
        raise NotImplementedError("lora_down/lora_act_out are only wired for the fused quantize epilogue (as in the reference's callers)")
    _require(rotary_emb is None or (norm_q is not None and norm_k is not None), "rotary_emb needs norm_q and norm_k (launch_impl:348-349)")
    alpha = 1.0 if alpha is None else float(alpha)
    _require(fp4 or alpha == 1.0, "INT4 requires alpha == 1 (gemm_w4a4_launch_impl.cuh:107)")
    _require(lora_up is None or lora_act_in is not None, "lora_up and lora_act_in go together (launch_impl:198)")

    K = act.shape[-1] * 2
    N = wgt.shape[0]
    _require(wgt.shape[1] * 2 == K, "act and wgt disagree on K")
    if out is not None:
        dtype = out.dtype
    elif not fp4:
        dtype = ascales.dtype            # gemm_w4a4.cu:63-73: INT4 infers the 16-bit type from ascales
    else:                                # NVFP4: from whichever 16-bit tensor is present (gemm_w4a4.cu:66-72)
        cands = [t for t in (bias, lora_up, lora_down, wcscales, norm_q, smooth_factor) if t is not None]
        _require(len(cands) > 0, "NVFP4 without out: pass bias, lora_up, wcscales or smooth_factor so that the 16-bit type is known")
        dtype = cands[0].dtype

    cs = None
    if wcscales is not None:
        cs = repack.channel_vector(wcscales, out_f32=True, mul=alpha)
    elif alpha != 1.0:
        cs = _const_vector(N, alpha, act.device)
    rank = 0 if lora_up is None else lora_up.shape[1]
    c = _Call(act=act, ascales=ascales, N=N, K=K, fp4=fp4, dtype=dtype, wgt=repack.qweight(wgt, fp4), wscales=repack.wscales(wscales, N, K, fp4),
              bias=None if bias is None else repack.channel_vector(bias, out_f32=True), cscale=cs, lora_act_in=lora_act_in,
              lora_up=repack.lora_up(lora_up, cs) if rank > 0 else None, rank=rank, lora_scales=lora_scales, out=out, act_unsigned=act_unsigned,
              mid_act=_mid_act(fuse_silu, fuse_gelu), norm_q=norm_q, norm_k=norm_k, rotary_emb=rotary_emb,
              out_qkv=(out_q, out_k, out_v) if pack_qkv else None, attn_tokens=attn_tokens)
    if litela_fused:
        _require(out.shape[0] == act.numel() // act.shape[-1], "out_linearattn: batch * tokens must equal the padded row count of act")
        c.out_vk, c.vk_tokens = out_vk, out_linearattn.shape[1]
    if fused_quant:
        c.qout, c.oscales = qout, oscales
        c.smooth_next = repack.channel_vector(smooth_factor, out_f32=False)
        if lora_down is not None and lora_down.shape[1] > 0:
            c.lora_down_next, c.lora_act_out, c.rank_down = repack.lora_down_next(lora_down), lora_act_out, lora_down.shape[1]
    _launch(c)
    if litela and not litela_fused:
        Mp = act.numel() // act.shape[-1]
        _require(out.shape[0] == Mp, "out_linearattn: batch * tokens must equal the padded row count of act")
        check(lib.nb200_litela_vk(torch_dtype_code(out.dtype), out.data_ptr(), out_linearattn.data_ptr(), out_vk.data_ptr(),
                                  out_linearattn.shape[0], out_linearattn.shape[1], N, torch.cuda.current_stream().cuda_stream), "litela_vk")


@on_device_of("q")
def linearattn_vk_mul_q(q: torch.Tensor, vk: torch.Tensor) -> None:
    """``kernels::linearattn_vk_mul_q`` (src/kernels/zgemm/gemm_w4a4.cu:107-111): in place on ``q`` [B, tokens, heads * 32]
    (or [B, tokens, heads, 32]) with ``vk`` float32 [B, heads, 33, 32]; eps = 1e-6 as in the reference launcher."""
    if not (q.is_cuda and vk.is_cuda):
        raise RuntimeError("nunchaku_b200 has no CPU path")
    if vk.dtype != torch.float32 or vk.dim() != 4 or vk.shape[2:] != (33, 32) or not vk.is_contiguous():
        raise ValueError("vk must be contiguous float32 [B, heads, 33, 32]")
    B, heads = vk.shape[0], vk.shape[1]
    if not q.is_contiguous() or q.shape[0] != B or q.numel() != B * q.shape[1] * heads * 32:
        raise ValueError("q must be contiguous [B, tokens, heads * 32]")
    check(lib.nb200_linearattn_vk_mul_q(torch_dtype_code(q.dtype), q.data_ptr(), vk.data_ptr(), B, q.shape[1], heads, 1e-6,
                                        torch.cuda.current_stream().cuda_stream), "linearattn_vk_mul_q")


_ws_cache: dict[tuple, torch.Tensor] = {}


def _reduce_workspace(Mp: int, rank: int, device) -> torch.Tensor:
    """Scratch of the fused quantise epilogue's deterministic reduction (include/nunchaku_b200.h: nb200_gemm_args.workspace), one per
    (device, stream, size)."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream, Mp, rank)
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.empty(int(lib.nb200_gemm_workspace_bytes(Mp, rank)), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


_const_cache: dict[tuple, torch.Tensor] = {}


def _const_vector(n: int, value: float, device) -> torch.Tensor:
    key = (n, value, str(device))
    t = _const_cache.get(key)
    if t is None:
        t = torch.full((n,), value, dtype=torch.float32, device=device)
        _const_cache[key] = t
    return t
