"""TEST INFRASTRUCTURE (oracle) -- the reference's AWQ W4A16 GEMV, restated on the CPU.

Reference: ``gemv_awq`` / ``gemv_kernel`` in src/kernels/awq/gemv_awq.cu:101-294 (called by GEMV_AWQ::forward, src/Linear.cpp:56-86,
for the AdaLN modulation linears).  Pinned against outputs of the reference kernel run on a B200 (tests/golden/ref_gpu_golden.npz,
keys ``awq*``; tests/test_ref_gpu_golden.py).

Weight layout (gemv_awq.cu:143-205): ``qweight`` int32 [OC/4, IC/8*4].  Output channels come in blocks of 8 = 2 groups ("idx") of
4 interleaved rows; a group is stored as [IC/64][row 4][64 k] 4-bit codes, 8 per u32, and inside each run of 32 k-elements (4 u32
w0..w3) element ``8*ii + 2*jj + e`` is nibble ``ii + 4*e`` of ``w_jj`` (the FasterTransformer i4 -> f16 conversion order followed by the
kernel's own shuffle).  Codes are unsigned 0..15; ``w = code * scale + zero`` (half FMA; ``zeros`` already holds -zero*scale), the
product ``w * x`` is rounded to the 16-bit type and accumulated in fp32 (gemv_awq.cu:222-236), the sum is rounded once.
"""
from __future__ import annotations

import torch

from .svdq import rn


def unpack_awq_qweight(qweight: torch.Tensor, OC: int, IC: int) -> torch.Tensor:
    """int32 [OC/4, IC/8*4] -> unsigned codes uint8 [OC, IC]"""
    assert qweight.numel() * 8 == OC * IC and OC % 8 == 0 and IC % 64 == 0
    w = qweight.reshape(-1).to(torch.int64) & 0xFFFFFFFF
    oc = torch.arange(OC).view(OC, 1)
    k = torch.arange(IC).view(1, IC)
    b, idx, r = oc // 8, (oc % 8) // 4, oc % 4
    pos = (k // 64) * 256 + r * 64 + (k % 64)                 # element position inside the 4-row group
    y = k % 32
    ii, jj, e = y // 8, (y % 8) // 2, y % 2
    word = b * IC + (idx * 4 * IC + (pos // 32) * 32) // 8 + jj
    nib = ii + 4 * e
    return ((w[word] >> (4 * nib)) & 0xF).to(torch.uint8)


def pack_awq_qweight(codes: torch.Tensor) -> torch.Tensor:
    """inverse of unpack_awq_qweight (test helper)"""
    OC, IC = codes.shape
    oc = torch.arange(OC).view(OC, 1)
    k = torch.arange(IC).view(1, IC)
    b, idx, r = oc // 8, (oc % 8) // 4, oc % 4
    pos = (k // 64) * 256 + r * 64 + (k % 64)
    y = k % 32
    ii, jj, e = y // 8, (y % 8) // 2, y % 2
    word = (b * IC + (idx * 4 * IC + (pos // 32) * 32) // 8 + jj).expand(OC, IC)
    vals = codes.to(torch.int64) << (4 * (ii + 4 * e)).expand(OC, IC)
    out = torch.zeros(OC * IC // 8, dtype=torch.int64)
    out.scatter_add_(0, word.reshape(-1), vals.reshape(-1))
    out = torch.where(out >= 2 ** 31, out - 2 ** 32, out)
    return out.to(torch.int32).view(OC // 4, IC // 8 * 4)


def gemv_awq(x: torch.Tensor, qweight: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, group_size: int = 64) -> torch.Tensor:
    """x [M, IC] hT, scales / zeros [IC/G, OC] hT -> [M, OC] hT, reference arithmetic"""
    hT = x.dtype
    G, OC = scales.shape
    IC = x.shape[1]
    q = unpack_awq_qweight(qweight, OC, IC).to(torch.float64)                                  # [OC, IC]
    s = scales.double().t().repeat_interleave(group_size, dim=1)                                # [OC, IC]
    z = zeros.double().t().repeat_interleave(group_size, dim=1)
    w = rn(q * s + z, hT).double()                                                              # __hfma2
    prod = rn(w.unsqueeze(0) * x.double().unsqueeze(1), hT)                                     # __hmul2, [M, OC, IC]
    acc = prod.to(torch.float32).double().sum(-1)                                               # fp32 accumulation (order unspecified)
    return rn(acc, hT)
