"""A layer's parameters in the layouts the B200 kernels consume (include/nunchaku_b200.h).

The reference hands its checkpoint tensors (mma.sync fragment order, nunchaku/lora/flux/packer.py) straight to its
kernels (src/Linear.cpp:124-154).  Here a layer is converted ONCE into a ``B200Weights`` bundle that the module owns:
no per-call cache lookups, an explicit lifetime (``SVDQW4A4Linear.invalidate()`` after a load / LoRA update), and the
checkpoint-layout copy can be dropped afterwards (``SVDQW4A4Linear.release_reference_layout()``), which halves the
resident weight memory.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, fields

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
        hts = [t.dtype for t in (proj_up, bias, smooth, wcscales, None if fp4 else wscales) if t is not None and t.dtype in (torch.float16, torch.bfloat16)]
        if not hts:
            raise ValueError("cannot tell the layer's 16-bit type: pass at least one of proj_up / bias / smooth / wcscales (gemm_w4a4.cu:63-73)")
        return cls(
            N=N, K=K, rank=rank, fp4=fp4, dtype=hts[0],
            qweight=repack.qweight(qweight, fp4, cache=False),
            wscales=repack.wscales(wscales, N, K, fp4, cache=False),
            bias=None if bias is None else repack.channel_vector(bias, out_f32=True, cache=False),
            cscale=cs,
            lora_up=repack.lora_up(proj_up, cs, cache=False) if rank > 0 else None,
            lora_down=repack.lora_down(proj_down, cache=False) if rank > 0 else None,
            lora_down_next=repack.lora_down_next(proj_down, cache=False) if rank > 0 else None,
            smooth=None if smooth is None else repack.channel_vector(smooth, out_f32=False, cache=False),
        )

    # ---- side-car persistence ------------------------------------------------------------------
    # The conversion is a pure function of the checkpoint tensors, so a deployment can do it once, keep the result next to the
    # checkpoint and skip both the repack kernels and the checkpoint-layout copy on every later start.  The file is a plain
    # ``torch.save`` dict of CPU tensors + geometry + a fingerprint of the source tensors it was made from.
    FORMAT = 1

    def state(self) -> dict:
        out = {"format": self.FORMAT, "N": self.N, "K": self.K, "rank": self.rank, "fp4": self.fp4, "dtype": str(self.dtype).replace("torch.", "")}
        for f in fields(self):
            v = getattr(self, f.name)
            if isinstance(v, torch.Tensor):
                out["t." + f.name] = v.detach().cpu()
        return out

    @classmethod
    def from_state(cls, state: dict, device) -> "B200Weights":
        if state.get("format") != cls.FORMAT:
            raise ValueError(f"unsupported B200Weights side-car format {state.get('format')!r} (this build reads {cls.FORMAT})")
        kw = {"N": int(state["N"]), "K": int(state["K"]), "rank": int(state["rank"]), "fp4": bool(state["fp4"]), "dtype": getattr(torch, state["dtype"])}
        for f in fields(cls):
            if f.name in kw:
                continue
            t = state.get("t." + f.name)
            kw[f.name] = None if t is None else t.to(device)
        w = cls(**kw)
        if w.qweight is None or tuple(w.qweight.shape) != (w.N, w.K // 2):
            raise ValueError("corrupt B200Weights side-car: qweight shape does not match its geometry")
        return w

    @staticmethod
    def fingerprint(**tensors) -> str:
        """sha256 over the raw bytes of the checkpoint tensors (and alpha) a bundle was converted from: a side-car is only reused for
        the checkpoint it was made from."""
        h = hashlib.sha256()
        for name in sorted(tensors):
            v = tensors[name]
            h.update(name.encode())
            if isinstance(v, torch.Tensor):
                h.update(str(tuple(v.shape)).encode() + str(v.dtype).encode())
                h.update(v.detach().contiguous().cpu().view(torch.uint8).numpy().tobytes())
            else:
                h.update(repr(v).encode())
        return h.hexdigest()

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.qweight, self.wscales, self.bias, self.cscale, self.lora_up, self.lora_down,
                                                          self.lora_down_next, self.smooth) if t is not None)
