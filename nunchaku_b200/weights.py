"""A layer's parameters in the layouts the B200 kernels consume (include/nunchaku_b200.h).

The reference hands its checkpoint tensors (mma.sync fragment order, nunchaku/lora/flux/packer.py) straight to its
kernels (src/Linear.cpp:124-154).  Here a layer is converted ONCE into a ``B200Weights`` bundle that the module owns:
no per-call cache lookups, an explicit lifetime (``SVDQW4A4Linear.invalidate()`` after a load / LoRA update), and the
checkpoint-layout copy can be dropped afterwards (``SVDQW4A4Linear.release_reference_layout()``), which halves the
resident weight memory.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import repack


@dataclass
class B200Weights:
    N: int
    K: int
    rank: int
    fp4: bool
    dtype: torch.dtype
    qweight: torch.Tensor                 # u8 [N, K/2], B200 nibble order
    wscales: torch.Tensor                 # INT4 hT [K/64, N] | NVFP4 ue4m3 tcgen05.cp tiles
    bias: torch.Tensor | None             # f32 [N]
    cscale: torch.Tensor | None           # f32 [N] = alpha * wcscales (None == 1)
    lora_up: torch.Tensor | None          # UMMA K-major blocks, pre-divided by cscale
    lora_down: torch.Tensor | None        # quantizer fragment order (this layer's input side)
    lora_down_next: torch.Tensor | None   # [R, K] row-major: TMA source when this layer is the NEXT layer of a fused fc1 epilogue
    smooth: torch.Tensor | None           # hT [K], natural order

    @classmethod
    def from_reference(cls, *, qweight, wscales, bias, smooth, proj_down, proj_up, wcscales=None, alpha: float | None = 1.0,
                       fp4: bool = False) -> "B200Weights":
        """Convert checkpoint-layout tensors (device) with the nb200_repack_* kernels; the result owns its storage."""
        N, K = qweight.shape[0], qweight.shape[1] * 2
        alpha = 1.0 if alpha is None else float(alpha)
        if not fp4 and alpha != 1.0:
            raise ValueError("INT4 requires alpha == 1 (gemm_w4a4_launch_impl.cuh:107)")
        cs = None
        if wcscales is not None and wcscales.numel() > 0:
            cs = repack.channel_vector(wcscales, out_f32=True, mul=alpha, cache=False)
        elif alpha != 1.0:
            cs = torch.full((N,), alpha, dtype=torch.float32, device=qweight.device)
        rank = 0 if proj_up is None else proj_up.shape[1]
        return cls(
            N=N, K=K, rank=rank, fp4=fp4, dtype=(proj_up if proj_up is not None else bias).dtype,
            qweight=repack.qweight(qweight, fp4, cache=False),
            wscales=repack.wscales(wscales, N, K, fp4, cache=False),
            bias=None if bias is None else repack.channel_vector(bias, out_f32=True, cache=False),
            cscale=cs,
            lora_up=repack.lora_up(proj_up, cs, cache=False) if rank > 0 else None,
            lora_down=repack.lora_down(proj_down, cache=False) if rank > 0 else None,
            lora_down_next=repack.lora_down_next(proj_down, cache=False) if rank > 0 else None,
            smooth=None if smooth is None else repack.channel_vector(smooth, out_f32=False, cache=False),
        )

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.qweight, self.wscales, self.bias, self.cscale, self.lora_up, self.lora_down,
                                                          self.lora_down_next, self.smooth) if t is not None)
