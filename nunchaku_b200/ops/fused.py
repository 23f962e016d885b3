"""The reference's fused operator entry points (nunchaku/ops/fused.py: ``fused_gelu_mlp``, ``fused_qkv_norm_rottary``) as thin
adapters over ``SVDQW4A4Linear.forward_mlp`` / ``forward_qkv`` (nunchaku_b200/models/linear.py), plus the one decision that is
new here: WHERE the MLP's intermediate goes on B200.

Measured: the quantise epilogue is tied to 128-wide tiles (NVFP4: shared-memory budget) whose operand
traffic per FLOP is a third higher than the 256-wide tiles'; above ~2k rows "plain GEMM with GELU on 256-wide tiles + the
activation quantizer" beats the fused launch even though the [M, hidden] tensor makes a round trip through L2 / HBM.  Same
arithmetic either way (GELU on the hT-rounded value, x / smooth, 4-bit codes + scales, low-rank down projection on the
unshifted GELU output); INT4 takes the two-launch route through the quantizer's ``shift_unsigned`` mode, which reproduces
the epilogue's shifted UNSIGNED codes (fc2 is built with ``act_unsigned=True`` either way).
"""
from __future__ import annotations

import torch

# None = choose by shape (below); True / False = force the fused / split route
FUSE_FC1_EPILOGUE: bool | None = None


def _fuse_fc1(fc1, rows: int) -> bool:
    if FUSE_FC1_EPILOGUE is not None:
        return bool(FUSE_FC1_EPILOGUE)
    return not (rows >= 2048 and fc1.out_features % 256 == 0)


def fused_gelu_mlp(x: torch.Tensor, fc1, fc2, pad_size: int = 256) -> torch.Tensor:
    """[B, S, C] -> fc2(gelu(fc1(x))) with fc2's 4-bit input produced without a separate 16-bit round trip where that wins."""
    b, s, c = x.shape
    y = fc1.forward_mlp(x.reshape(b * s, c), fc2, fuse=_fuse_fc1(fc1, b * s), pad_size=pad_size)
    return y.view(b, s, fc2.out_features)


def fused_qkv_norm_rottary(x: torch.Tensor, proj, norm_q=None, norm_k=None, rotary_emb: torch.Tensor | None = None,
                           output: torch.Tensor | tuple | None = None, attn_tokens: int = 0):
    """QKV projection + per-head RMSNorm(Q, K) + RoPE in one launch.  ``norm_q`` / ``norm_k``: modules with a 128-element
    ``weight`` (torch.nn.RMSNorm in the reference); ``output``: a [B*S, 3*H*128] tensor, or a tuple of three fp16
    [1, H, tokens_pad, 128] tensors for the attention-ready layout."""
    b, s, c = x.shape
    if norm_q is None or norm_k is None or rotary_emb is None:
        raise ValueError("fused_qkv_norm_rottary needs norm_q, norm_k and rotary_emb")
    x2d = x.reshape(b * s, c)
    if isinstance(output, tuple):
        if len(output) != 3:
            raise ValueError("output tuple must be (out_q, out_k, out_v)")
        return proj.forward_qkv(x2d, norm_q.weight, norm_k.weight, rotary_emb, out_qkv=output, attn_tokens=attn_tokens)
    y = proj.forward_qkv(x2d, norm_q.weight, norm_k.weight, rotary_emb, output=None if output is None else output.view(b * s, -1))
    return y.view(b, s, proj.out_features)
