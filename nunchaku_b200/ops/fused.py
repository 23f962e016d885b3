"""Mirror of nunchaku/ops/fused.py (fused_gelu_mlp :14-79, fused_qkv_norm_rottary :82-178)."""
from __future__ import annotations

import torch

from ..utils import ceil_divide
from .gemm import svdq_gemm_w4a4_cuda
from .quantize import svdq_quantize_w4a4_act_fuse_lora_cuda


# NVFP4, large M: where the intermediate goes.  None = choose by shape, True / False = force.  Measured on B200
# (DESIGN.md section 4.4): the quantise epilogue is tied to 128-wide tiles (shared-memory budget), whose operand traffic
# per FLOP is 1/3 higher than the plain 256-wide kernel's; above ~2k rows "plain GEMM with GELU (256-wide) + the
# activation quantizer" is faster than the fused launch even though the [M, hidden] tensor makes a round trip through
# L2 / HBM.  Same arithmetic either way (GELU on the hT-rounded value, x / smooth, e2m1 codes + ue4m3 scales, LoRA-down
# on the unshifted GELU output).  INT4 takes the same two-launch route through the quantizer's ``shift_unsigned`` mode, which
# reproduces the epilogue's shifted UNSIGNED codes (fc2 is built with act_unsigned=True either way).
FUSE_FC1_EPILOGUE: bool | None = None


def _fuse_fc1(fc1, rows: int) -> bool:
    if FUSE_FC1_EPILOGUE is not None:
        return bool(FUSE_FC1_EPILOGUE)
    return not (rows >= 2048 and fc1.out_features % 256 == 0)


def fused_gelu_mlp(x: torch.Tensor, fc1, fc2, pad_size: int = 256) -> torch.Tensor:
    """fc1 -> GELU -> fc2 where fc1's GEMM epilogue already produces fc2's 4-bit input, its scales
    and its low-rank hidden state; the [M, hidden] 16-bit tensor never touches HBM.
    INT4: the GELU output is shifted by 0.171875 and quantised UNSIGNED, so ``fc2`` must have been
    built with ``act_unsigned=True`` (its bias absorbs the shift), exactly as in the reference."""
    batch_size, seq_len, channels = x.shape
    x = x.view(batch_size * seq_len, channels)
    quantized_x, ascales, lora_act = fc1.quantize(x)
    if not _fuse_fc1(fc1, batch_size * seq_len):
        hidden = torch.empty(batch_size * seq_len, fc1.out_features, dtype=x.dtype, device=x.device)
        fp4 = fc1.precision == "nvfp4"
        svdq_gemm_w4a4_cuda(act=quantized_x, wgt=fc1.qweight, out=hidden, ascales=ascales, wscales=fc1.wscales, lora_act_in=lora_act,
                            lora_up=fc1.proj_up, bias=fc1.bias, fp4=fp4, alpha=fc1.wtscale, wcscales=fc1.wcscales, fuse_gelu=True)
        q2, s2, la2 = svdq_quantize_w4a4_act_fuse_lora_cuda(hidden, lora_down=fc2.proj_down, smooth=fc2.smooth_factor, fp4=fp4,
                                                            pad_size=pad_size, shift_unsigned=not fp4)
        output = torch.empty(batch_size * seq_len, fc2.out_features, dtype=x.dtype, device=x.device)
        return fc2.forward_quant(q2, s2, la2, output=output).view(batch_size, seq_len, -1)
    batch_size_pad = ceil_divide(batch_size * seq_len, pad_size) * pad_size
    qout_act = torch.empty(batch_size_pad, fc1.out_features // 2, dtype=torch.uint8, device=x.device)
    if fc2.precision == "nvfp4":
        qout_ascales = torch.empty(fc1.out_features // 16, batch_size_pad, dtype=torch.float8_e4m3fn, device=x.device)
    else:
        qout_ascales = torch.empty(fc1.out_features // 64, batch_size_pad, dtype=x.dtype, device=x.device)
    qout_lora_act = torch.empty(batch_size_pad, fc2.proj_down.shape[1], dtype=torch.float32, device=x.device)
    svdq_gemm_w4a4_cuda(
        act=quantized_x,
        wgt=fc1.qweight,
        qout=qout_act,
        ascales=ascales,
        wscales=fc1.wscales,
        oscales=qout_ascales,
        lora_act_in=lora_act,
        lora_up=fc1.proj_up,
        lora_down=fc2.proj_down,
        lora_act_out=qout_lora_act,
        bias=fc1.bias,
        smooth_factor=fc2.smooth_factor,
        fp4=fc1.precision == "nvfp4",
        alpha=fc1.wtscale,
        wcscales=fc1.wcscales,
    )
    output = torch.empty(batch_size * seq_len, fc2.out_features, dtype=x.dtype, device=x.device)
    output = fc2.forward_quant(qout_act, qout_ascales, qout_lora_act, output=output)
    return output.view(batch_size, seq_len, -1)


def fused_qkv_norm_rottary(
    x: torch.Tensor,
    proj,
    norm_q=None,
    norm_k=None,
    rotary_emb: torch.Tensor | None = None,
    output: torch.Tensor | tuple | None = None,
    attn_tokens: int = 0,
):
    """QKV projection with per-head RMSNorm on Q/K and rotary embedding fused in the GEMM epilogue.
    ``norm_q``/``norm_k`` are modules with a ``weight`` of 128 elements (torch.nn.RMSNorm in the
    reference); ``rotary_emb`` is the reference's *packed* table (``pack_rotemb``)."""
    batch_size, seq_len, channels = x.shape
    x = x.view(batch_size * seq_len, channels)
    quantized_x, ascales, lora_act = proj.quantize(x)
    if isinstance(output, tuple):
        # attention-ready fp16 Q / K / V, [1, heads, tokens_pad, 128] each (nunchaku/ops/fused.py:137-159)
        assert len(output) == 3
        output_q, output_k, output_v = output
        svdq_gemm_w4a4_cuda(
            act=quantized_x,
            wgt=proj.qweight,
            ascales=ascales,
            wscales=proj.wscales,
            lora_act_in=lora_act,
            lora_up=proj.proj_up,
            bias=proj.bias,
            fp4=proj.precision == "nvfp4",
            alpha=proj.wtscale,
            wcscales=proj.wcscales,
            norm_q=norm_q.weight if norm_q is not None else None,
            norm_k=norm_k.weight if norm_k is not None else None,
            rotary_emb=rotary_emb,
            out_q=output_q,
            out_k=output_k,
            out_v=output_v,
            attn_tokens=attn_tokens,
        )
        return output_q, output_k, output_v
    if output is None:
        output = torch.empty(batch_size * seq_len, proj.out_features, dtype=x.dtype, device=x.device)
    svdq_gemm_w4a4_cuda(
        act=quantized_x,
        wgt=proj.qweight,
        out=output,
        ascales=ascales,
        wscales=proj.wscales,
        lora_act_in=lora_act,
        lora_up=proj.proj_up,
        bias=proj.bias,
        fp4=proj.precision == "nvfp4",
        alpha=proj.wtscale,
        wcscales=proj.wcscales,
        norm_q=norm_q.weight if norm_q is not None else None,
        norm_k=norm_k.weight if norm_k is not None else None,
        rotary_emb=rotary_emb,
    )
    return output.view(batch_size, seq_len, -1)
