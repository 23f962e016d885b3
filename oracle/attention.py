"""CPU oracle of the reference's attention_fp16 (src/kernels/zgemm/attention.cu:10-94, attention.cuh) -- TEST INFRASTRUCTURE, never
imported by the product path.

Exact-math restatement (fp64): softmax in base 2 with scale * log2(e), NaN scores (K's pad rows, attention.cuh:192-221) masked to -inf.
The reference accumulates Q K^T and P V in fp16 per 32-key tile and rounds P to fp16, so it sits ~1e-3 (fp16 output) from this; the pinned
distance is in tests/golden/ref_gpu_golden_report.json (``attn_*``: reference kernel on a B200 vs fp64) and
tests/test_ref_gpu_golden.py::test_attention_oracle_vs_reference_gpu."""
from __future__ import annotations

import torch


def pack_qkv_rowmajor(qkv: torch.Tensor, heads: int, tokens_pad: int):
    """[T, 3 * H * 128] hT -> (q, k, v) fp16 [1, H, T_pad, 128] in the B200 PackQKV layout: row-major inside a head, pad rows 0 / NaN / 0
    (what nb200_gemm_w4a4 writes into out_q / out_k / out_v; masks as the reference's EpiloguePackQKV, epilogues.cuh:446-470)."""
    T = qkv.shape[0]
    parts = []
    for i, fill in enumerate((0.0, float("nan"), 0.0)):
        t = torch.full((1, heads, tokens_pad, 128), fill, dtype=torch.float16)
        t[0, :, :T] = qkv[:, i * heads * 128:(i + 1) * heads * 128].float().to(torch.float16).view(T, heads, 128).transpose(0, 1)
        parts.append(t)
    return tuple(parts)


def attention_fp16(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float) -> torch.Tensor:
    """q / k / v [B, H, T, 128] (any float dtype, NaN rows of k = masked keys) -> fp64 [B, T_q, H * 128]"""
    qd, kd, vd = q.double(), k.double(), v.double()
    s = torch.einsum("bhid,bhjd->bhij", qd, kd) * (scale * 1.4426950408889634)
    s = torch.where(torch.isnan(s), torch.full_like(s, float("-inf")), s)
    m = s.amax(-1, keepdim=True)
    p = torch.exp2(s - m)
    vd = torch.where(torch.isnan(vd), torch.zeros_like(vd), vd)
    o = torch.einsum("bhij,bhjd->bhid", p, vd) / p.sum(-1, keepdim=True)
    B, H, T, D = o.shape
    return o.permute(0, 2, 1, 3).reshape(B, T, H * D)
