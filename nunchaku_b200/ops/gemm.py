"""Mirror of nunchaku/ops/gemm.py:12-160 (svdq_gemm_w4a4_cuda) on B200."""
from __future__ import annotations

import ctypes
import math

import torch

from .. import repack
from .._C import NB200_ACT_GELU, NB200_ACT_NONE, NB200_ACT_SILU, NB200_MAX_LORA_SCALES, GemmArgs, check, lib
from ..utils import on_device_of, torch_dtype_code

# tuning knobs for experiments (0 = let the launcher choose)
BLOCK_N_OVERRIDE = 0
NUM_SMS_OVERRIDE = 0
PROF_BUFFER = None   # torch int64 tensor [grid, 16]: per-CTA barrier-wait cycle counters (tools/gemm_prof.py)


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
):
    """Fused SVDQuant W4A4 GEMM; positional/keyword compatible with the reference wrapper
    (nunchaku/ops/gemm.py:12-160).  Results are written in place into the provided outputs.
    ``fuse_gelu`` (keyword-only, not in the reference signature) applies the reference's tanh-GELU in the plain
    epilogue (C ABI ``mid_act = NB200_ACT_GELU``; the reference only reaches GELU through the fused quantise epilogue).

    ``act`` / ``ascales`` / ``lora_act_in`` must come from this package's quantize op (or fused
    epilogue); ``wgt`` / ``wscales`` / ``lora_up`` / ``bias`` / ``wcscales`` are the reference's
    packed checkpoint tensors and are repacked once on first use (nunchaku_b200.repack).
    ``poolout`` is accepted and ignored exactly as in the reference.
    """
    if act is None or wgt is None or ascales is None or wscales is None:
        raise ValueError("act, wgt, ascales and wscales are required")
    if not act.is_cuda:
        raise RuntimeError("nunchaku_b200 has no CPU path: act must be a CUDA tensor")
    litela = out_vk is not None or out_linearattn is not None
    if litela:
        # SANA linear attention (launch_impl:311-346): relu(Q) -> out_linearattn, per-head V^T relu(K) state -> out_vk.
        # The GEMM writes the plain hT projection into a scratch tensor and nb200_litela_vk reduces it while it is L2-hot.
        if out_vk is None or out_linearattn is None:
            raise ValueError("out_vk and out_linearattn go together (launch_impl:313)")
        if out_vk.dtype != torch.float32 or out_vk.dim() != 4 or out_vk.shape[2:] != (33, 32):
            raise ValueError("out_vk must be float32 [B, heads, 33, 32]")
        if out_linearattn.dim() != 3 or out_linearattn.shape[0] != out_vk.shape[0]:
            raise ValueError("out_linearattn must be [B, tokens, N / 3]")
        if not (out_vk.is_contiguous() and out_linearattn.is_contiguous()):
            raise ValueError("out_vk / out_linearattn must be contiguous")
        if out_linearattn.shape[1] % 256 != 0:
            raise ValueError("tokens must be a multiple of 256 (launch_impl:331)")
        if out_vk.shape[1] * 96 != wgt.shape[0] or out_linearattn.shape[2] * 3 != wgt.shape[0]:
            raise ValueError("N must be 3 * heads * 32")
        out = torch.empty(out_linearattn.shape[0] * out_linearattn.shape[1], wgt.shape[0], dtype=out_linearattn.dtype, device=act.device)
    pack_qkv = out_q is not None or out_k is not None or out_v is not None
    if pack_qkv:
        # EpiloguePackQKV (launch_impl:376-393): fp16 [B=1, heads, rows >= Mp, 128], plain row-major inside a head
        if out_q is None or out_k is None or out_v is None:
            raise ValueError("out_q, out_k and out_v go together")
        if rotary_emb is None:
            raise ValueError("out_q/out_k/out_v need rotary_emb, norm_q and norm_k (launch_impl:347-376)")
        for t in (out_q, out_k, out_v):
            if t.dtype != torch.float16 or t.dim() != 4 or t.shape[0] != 1 or t.shape[-1] != 128:
                raise ValueError("out_q/out_k/out_v must be float16 [1, heads, tokens_pad, 128]")
            if t.stride(-1) != 1 or t.stride(-2) != 128:
                raise ValueError("out_q/out_k/out_v: the last two dims must be contiguous")
    fused_quant = qout is not None and oscales is not None          # launch_impl:282
    if (qout is None) != (oscales is None):
        raise ValueError("qout and oscales go together")
    if out is None and not fused_quant and not pack_qkv:
        raise ValueError("out is required unless qout/oscales or out_q/out_k/out_v are given")
    if pack_qkv:
        out = None  # the reference ignores `out` in this mode (launch_impl:376-393)
    if fused_quant and smooth_factor is None:
        raise ValueError("qout needs smooth_factor (the next layer's smoothing vector)")
    if (lora_down is None) != (lora_act_out is None):
        raise ValueError("lora_down and lora_act_out go together (launch_impl:199)")
    if lora_down is not None and not fused_quant:
        raise NotImplementedError("lora_down/lora_act_out are only wired for the fused quantize epilogue (as in the reference's callers)")
    if rotary_emb is not None and (norm_q is None or norm_k is None):
        raise ValueError("rotary_emb needs norm_q and norm_k (launch_impl:348-349)")
    if alpha is None:
        alpha = 1.0
    if not fp4 and float(alpha) != 1.0:
        raise ValueError("INT4 requires alpha == 1 (gemm_w4a4_launch_impl.cuh:107)")

    Mp = act.numel() // act.shape[-1]
    K = act.shape[-1] * 2
    N = wgt.shape[0]
    if wgt.shape[1] * 2 != K:
        raise ValueError("act and wgt disagree on K")
    out2d = None if out is None else out.view(-1, out.shape[-1])
    if out is not None:
        dtype = out.dtype
    elif not fp4:
        dtype = ascales.dtype            # gemm_w4a4.cu:63-73: INT4 infers the 16-bit type from ascales
    else:
        dtype = bias.dtype if bias is not None else (lora_up.dtype if lora_up is not None else norm_q.dtype)

    args = GemmArgs()
    args.act = act.data_ptr()
    args.wgt = repack.qweight(wgt, fp4).data_ptr()
    args.ascales = ascales.data_ptr()
    args.wscales = repack.wscales(wscales, N, K, fp4).data_ptr()
    cs = None
    if wcscales is not None or float(alpha) != 1.0:
        if wcscales is not None:
            cs = repack.channel_vector(wcscales, out_f32=True, mul=float(alpha))
        else:
            cs = _const_vector(N, float(alpha), act.device)
    args.cscale = None if cs is None else cs.data_ptr()
    args.bias = None if bias is None else repack.channel_vector(bias, out_f32=True).data_ptr()
    rank = 0
    if lora_up is not None:
        if lora_act_in is None:
            raise ValueError("lora_up and lora_act_in go together (launch_impl:198)")
        rank = lora_up.shape[1]
        if rank > 0:
            args.lora_up = repack.lora_up(lora_up, cs).data_ptr()
            args.lora_act_in = lora_act_in.data_ptr()
            assert lora_act_in.shape[-1] == rank and lora_act_in.dtype == torch.float32
    if lora_scales is None:
        lora_scales = [1.0] * math.ceil(rank / 16)
    for i in range(NB200_MAX_LORA_SCALES):  # absent groups get 0 (launch_impl:220-225)
        args.lora_scales[i] = float(lora_scales[i]) if i < len(lora_scales) else 0.0
    args.Mp, args.N, args.K = Mp, N, K
    if out2d is not None:
        args.out = out2d.data_ptr()
        args.M_out, args.N_out = out2d.shape[0], out2d.shape[1]
    args.R_up, args.R_down = rank, 0
    if fused_quant:
        assert qout.shape[-1] * 2 == N and qout.numel() // qout.shape[-1] == Mp and qout.is_contiguous()
        args.qout = qout.data_ptr()
        args.oscales = oscales.data_ptr()
        args.smooth_next = repack.channel_vector(smooth_factor, out_f32=False).data_ptr()
        if lora_down is not None and lora_down.shape[1] > 0:
            assert lora_act_out.dtype == torch.float32 and lora_act_out.shape == (Mp, lora_down.shape[1])
            args.lora_down_next = repack.lora_down_next(lora_down).data_ptr()
            args.lora_act_out = lora_act_out.data_ptr()
            args.R_down = lora_down.shape[1]
    if rotary_emb is not None:
        assert rotary_emb.dtype == torch.float32 and rotary_emb.numel() == Mp * 128, "rotary_emb: packed fp32 [Mp, 128]"
        assert norm_q.numel() == 128 and norm_k.numel() == 128
        args.rotary_emb = rotary_emb.data_ptr()
        args.norm_q = norm_q.contiguous().data_ptr()
        args.norm_k = norm_k.contiguous().data_ptr()
        if pack_qkv:
            heads = N // 384
            for t in (out_q, out_k, out_v):
                if t.shape[1] != heads or t.shape[2] < Mp:
                    raise ValueError(f"out_q/out_k/out_v must hold {heads} heads of >= {Mp} rows")
            args.out_q, args.out_k, args.out_v = out_q.data_ptr(), out_k.data_ptr(), out_v.data_ptr()
            args.stride_head_q, args.stride_head_k, args.stride_head_v = out_q.stride(1), out_k.stride(1), out_v.stride(1)
            args.attn_tokens = int(attn_tokens)
    args.dtype = torch_dtype_code(dtype)
    args.fp4 = int(fp4)
    args.act_unsigned = int(act_unsigned)
    if fuse_silu and fuse_gelu:
        raise ValueError("fuse_silu and fuse_gelu are exclusive")
    args.mid_act = NB200_ACT_SILU if fuse_silu else (NB200_ACT_GELU if fuse_gelu else NB200_ACT_NONE)
    args.block_n = BLOCK_N_OVERRIDE
    args.num_sms = NUM_SMS_OVERRIDE
    args.prof = None if PROF_BUFFER is None else PROF_BUFFER.data_ptr()
    check(lib.nb200_gemm_w4a4(ctypes.byref(args), torch.cuda.current_stream().cuda_stream), "gemm_w4a4")
    if litela:
        if out.shape[0] != Mp:
            raise ValueError("out_linearattn: batch * tokens must equal the padded row count of act")
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


_const_cache: dict[tuple, torch.Tensor] = {}


def _const_vector(n: int, value: float, device) -> torch.Tensor:
    key = (n, value, str(device))
    t = _const_cache.get(key)
    if t is None:
        t = torch.full((n,), value, dtype=torch.float32, device=device)
        _const_cache[key] = t
    return t
