"""``SVDQW4A4Linear`` for B200: state-dict compatible with the reference module (nunchaku/models/linear.py:13-274 -- same
constructor keywords, parameter names, shapes and dtypes, so its checkpoints load unchanged), built around a layer-owned
set of converted weights instead of per-call tensor plumbing:

* parameters are declared from one table (``_PARAMS``) and keep the checkpoint layout;
* ``b200()`` converts them once into a ``B200Weights`` bundle (the kernels' layouts); a ``load_state_dict`` hook and
  ``invalidate()`` drop it, ``release_reference_layout()`` frees the checkpoint-layout storage afterwards;
* the forward variants the reference spreads over ``forward`` / ``forward_quant`` / ``nunchaku.ops.fused`` are methods
  here: ``forward``, ``forward_quant``, ``forward_mlp`` (fc1 -> GELU -> fc2 with fc2's 4-bit input produced by fc1's
  epilogue, or by the two-launch split route on large M) and ``forward_qkv`` (RMSNorm + RoPE [+ PackQKV] epilogue).
"""
from __future__ import annotations

from dataclasses import fields as dataclass_fields

import torch
from torch import nn

from ..ops.gemm import gemm_b200
from ..ops.quantize import quantize_b200
from ..utils import ceil_divide
from ..weights import B200Weights

_GROUP = {"int4": 64, "nvfp4": 16}

# name -> (shape as a function of (in, out, rank, group), dtype selector, trainable)   [reference schema, linear.py:68-131]
_PARAMS = {
    "qweight": (lambda i, o, r, g: (o, i // 2), "int8", False),
    "wscales": (lambda i, o, r, g: (i // g, o), "scale", False),
    "smooth_factor": (lambda i, o, r, g: (i,), "half", False),
    "smooth_factor_orig": (lambda i, o, r, g: (i,), "half", False),
    "proj_down": (lambda i, o, r, g: (i, r), "half", True),
    "proj_up": (lambda i, o, r, g: (o, r), "half", True),
}


class SVDQW4A4Linear(nn.Module):
    """4-bit weights x 4-bit activations + a 16-bit rank-``rank`` branch, fused on tcgen05 (csrc/gemm_*.cu)."""

    def __init__(self, in_features: int, out_features: int, rank: int = 32, bias: bool = True, precision: str = "int4",
                 act_unsigned: bool = False, torch_dtype: torch.dtype = torch.bfloat16, device: str | torch.device | None = None):
        super().__init__()
        if precision not in _GROUP:
            raise ValueError(f"Invalid precision: {precision}")
        self.in_features, self.out_features, self.rank = in_features, out_features, rank
        self.precision, self.torch_dtype, self.act_unsigned = precision, torch_dtype, act_unsigned
        self.group_size = _GROUP[precision]
        device = torch.device("cpu") if device is None else device
        kinds = {"int8": torch.int8, "half": torch_dtype, "scale": torch_dtype if precision == "int4" else torch.float8_e4m3fn}
        for name, (shape, kind, trainable) in _PARAMS.items():
            t = torch.empty(shape(in_features, out_features, rank, self.group_size), dtype=kinds[kind], device=device)
            self.register_parameter(name, nn.Parameter(t, requires_grad=trainable))
        self.bias = nn.Parameter(torch.empty(out_features, dtype=torch_dtype, device=device)) if bias else None
        if precision == "nvfp4":    # per-channel and per-tensor weight scales of NVFP4 checkpoints (linear.py:124-131)
            self.wcscales = nn.Parameter(torch.ones(out_features, dtype=torch_dtype, device=device), requires_grad=False)
            self.wtscale = 1.0
        else:
            self.wcscales, self.wtscale = None, None
        self._b200: B200Weights | None = None
        self._b200_alpha = None
        self.register_load_state_dict_post_hook(lambda module, _keys: module.invalidate())

    @classmethod
    def from_linear(cls, linear: nn.Linear, **kwargs) -> "SVDQW4A4Linear":
        """An (uninitialised) quantised layer with the geometry, dtype and device of ``linear`` (linear.py:137-159)."""
        kwargs.setdefault("in_features", linear.in_features)
        return cls(out_features=linear.out_features, bias=linear.bias is not None, torch_dtype=linear.weight.dtype,
                   device=linear.weight.device, **kwargs)

    # ---- converted weights -----------------------------------------------------------------------------------------
    @property
    def fp4(self) -> bool:
        return self.precision == "nvfp4"

    def b200(self) -> B200Weights:
        """The layer in the kernels' layouts; converted on first use and after ``invalidate()`` (or a changed ``wtscale``)."""
        if self._b200 is None or self._b200_alpha != self.wtscale:
            if not self.qweight.is_cuda:
                raise RuntimeError("nunchaku_b200 has no CPU path: move the module to a CUDA device")
            if self.qweight.numel() == 0:
                raise RuntimeError("the checkpoint-layout parameters were released; reload the state dict to convert again")
            self._b200 = B200Weights.from_reference(qweight=self.qweight.data, wscales=self.wscales.data,
                                                    bias=None if self.bias is None else self.bias.data, smooth=self.smooth_factor.data,
                                                    proj_down=self.proj_down.data, proj_up=self.proj_up.data,
                                                    wcscales=None if self.wcscales is None else self.wcscales.data, alpha=self.wtscale, fp4=self.fp4)
            self._b200_alpha = self.wtscale
        return self._b200

    def _apply(self, fn, recurse=True):
        """``module.to(device)`` / ``.cuda(i)`` after the conversion: the converted bundle follows the parameters to their new device (it cannot be
        rebuilt once ``release_reference_layout()`` has dropped the checkpoint layout).  Dtype casts do not touch it (its fp32 vectors stay fp32)."""
        out = super()._apply(fn, recurse)
        w = self._b200
        if w is not None and w.qweight.device != self.qweight.device:
            dev = self.qweight.device
            for f in dataclass_fields(w):
                v = getattr(w, f.name)
                if isinstance(v, torch.Tensor):
                    setattr(w, f.name, v.to(dev))
        return out

    def invalidate(self) -> None:
        """Call after changing parameters in place (LoRA merge, ``param.data.copy_``); ``load_state_dict`` does it itself."""
        self._b200 = None

    def _source_fingerprint(self) -> str:
        return B200Weights.fingerprint(qweight=self.qweight.data, wscales=self.wscales.data, bias=None if self.bias is None else self.bias.data,
                                       smooth=self.smooth_factor.data, proj_down=self.proj_down.data, proj_up=self.proj_up.data,
                                       wcscales=None if self.wcscales is None else self.wcscales.data, alpha=self.wtscale, precision=self.precision)

    def save_b200(self, path: str) -> None:
        """Write the converted layer to a side-car file: ``load_b200`` on a later start skips the conversion, and -- with
        ``require_source=False`` on a module whose checkpoint-layout tensors were never loaded -- the checkpoint-layout copy as well."""
        if self.qweight.numel() == 0:
            raise RuntimeError("the checkpoint-layout parameters were released: save the side-car before release_reference_layout()")
        state = self.b200().state()
        state["source_sha256"] = self._source_fingerprint()
        state["precision"], state["act_unsigned"] = self.precision, bool(self.act_unsigned)
        torch.save(state, path)

    def load_b200(self, path: str, *, require_source: bool = True) -> None:
        """Adopt a side-car written by ``save_b200``.  ``require_source=True`` (default): the file must have been made from exactly the
        parameters this module holds now (sha256 of their bytes), otherwise ValueError.  ``require_source=False``: trust the file (geometry
        and precision are still checked) -- for deployments that ship only the side-car; the module's checkpoint-layout tensors are then
        released."""
        state = torch.load(path, map_location="cpu", weights_only=True)
        if state.get("precision") != self.precision or int(state["N"]) != self.out_features or int(state["K"]) != self.in_features or \
                int(state["rank"]) != self.rank:
            raise ValueError("B200 side-car does not match this layer's geometry / precision")
        if require_source:
            if self.qweight.numel() == 0 or state.get("source_sha256") != self._source_fingerprint():
                raise ValueError("B200 side-car was made from different parameters than this module holds")
        self._b200 = B200Weights.from_state(state, self.qweight.device)
        self._b200_alpha = self.wtscale
        if not require_source:
            for name in ("qweight", "wscales"):
                p = getattr(self, name)
                p.data = torch.empty(0, dtype=p.dtype, device=p.device)

    def release_reference_layout(self) -> int:
        """Convert, then free the checkpoint-layout storage of the big tensors (halves the resident weight memory; the
        reference keeps only its own layout).  The state dict of a released module is no longer loadable elsewhere.
        Returns the number of bytes freed."""
        self.b200()
        freed = 0
        for name in ("qweight", "wscales"):
            p = getattr(self, name)
            freed += p.numel() * p.element_size()
            p.data = torch.empty(0, dtype=p.dtype, device=p.device)
        return freed

    # ---- the reference's three methods -------------------------------------------------------------------------------
    def quantize(self, x: torch.Tensor, pad_size: int = 256, *, shift_unsigned: bool = False):
        """[M, in] hT -> (codes u8 [M_pad, in/2], scales [in/G, M_pad], low-rank hidden state f32 [M_pad, rank])."""
        return quantize_b200(x, self.b200(), pad_size=pad_size, shift_unsigned=shift_unsigned)

    def forward_quant(self, quantized_x, ascales, lora_act, output=None, *, fuse_gelu: bool = False, fuse_silu: bool = False):
        w = self.b200()
        if output is None:
            output = torch.empty(quantized_x.shape[0], self.out_features, dtype=w.dtype, device=quantized_x.device)
        gemm_b200(quantized_x, ascales, lora_act, w, out=output, act_unsigned=self.act_unsigned, fuse_gelu=fuse_gelu, fuse_silu=fuse_silu)
        return output

    def forward(self, x: torch.Tensor, output: torch.Tensor | None = None) -> torch.Tensor:
        """[B, S, in] -> [B, S, out]"""
        if x.dim() != 3:
            raise ValueError("expected a [batch, tokens, in_features] input (linear.py:181)")
        rows = x.shape[0] * x.shape[1]
        if output is None:
            output = torch.empty(rows, self.out_features, dtype=x.dtype, device=x.device)
        y = self.forward_quant(*self.quantize(x.reshape(rows, x.shape[2])), output.view(rows, self.out_features))
        return y.view(x.shape[0], x.shape[1], self.out_features)

    # ---- fused variants (nunchaku/ops/fused.py in the reference) -------------------------------------------------------
    def quantize_next(self, quantized_x, ascales, lora_act, nxt: "SVDQW4A4Linear", pad_size: int = 256):
        """This layer's GEMM with GELU, handing ``nxt`` its 4-bit input, scales and low-rank hidden state straight from the
        epilogue (the [M, out] tensor never reaches HBM).  INT4: the shifted GELU output is quantised UNSIGNED, so ``nxt``
        must have been built with ``act_unsigned=True`` (its bias absorbs the shift), exactly as in the reference."""
        w, wn = self.b200(), nxt.b200()
        Mp = quantized_x.shape[0]
        dev = quantized_x.device
        q2 = torch.empty(Mp, self.out_features // 2, dtype=torch.uint8, device=dev)
        s2 = torch.empty(self.out_features // nxt.group_size, Mp, dtype=torch.float8_e4m3fn if nxt.fp4 else w.dtype, device=dev)
        la2 = torch.empty(Mp, wn.rank, dtype=torch.float32, device=dev)
        gemm_b200(quantized_x, ascales, lora_act, w, act_unsigned=self.act_unsigned, next_w=wn, qout=q2, oscales=s2, lora_act_out=la2)
        return q2, s2, la2

    def forward_mlp(self, x2d: torch.Tensor, nxt: "SVDQW4A4Linear", *, fuse: bool, pad_size: int = 256) -> torch.Tensor:
        """fc1 (= self) -> GELU -> fc2 (= nxt) on a [M, in] input.  ``fuse`` = one launch for fc1 + the hand-off; otherwise
        fc1 with GELU in the plain epilogue followed by the activation quantizer (INT4: its ``shift_unsigned`` mode), which
        produces bit-identical hand-off tensors (tests/test_gpu_fused.py::test_fused_gelu_mlp_both_routes_agree)."""
        q, s, la = self.quantize(x2d, pad_size)
        if fuse:
            q2, s2, la2 = self.quantize_next(q, s, la, nxt, pad_size)
        else:
            hidden = self.forward_quant(q, s, la, fuse_gelu=True)[: x2d.shape[0]]
            q2, s2, la2 = nxt.quantize(hidden, pad_size, shift_unsigned=not nxt.fp4)
        return nxt.forward_quant(q2, s2, la2, torch.empty(x2d.shape[0], nxt.out_features, dtype=x2d.dtype, device=x2d.device))

    def forward_qkv(self, x2d: torch.Tensor, norm_q: torch.Tensor, norm_k: torch.Tensor, rotary_emb: torch.Tensor, *, output=None,
                    out_qkv: tuple | None = None, attn_tokens: int = 0):
        """QKV projection with per-head RMSNorm on Q / K and the rotary embedding applied in the GEMM epilogue; ``rotary_emb`` is
        the reference's packed table (``pack_rotemb``).  ``out_qkv`` = three fp16 [1, heads, tokens_pad, 128] tensors instead
        of the row-major [M, out] result (EpiloguePackQKV)."""
        w = self.b200()
        q, s, la = self.quantize(x2d)
        if out_qkv is not None:
            # NVFP4 at model sizes: a [Mp, N] scratch lets the launcher run the plain GEMM + the RMSNorm / RoPE / pack kernel (csrc/rope.cu)
            # instead of the fused epilogue -- same bits, faster (include/nunchaku_b200.h: out next to out_q/k/v)
            scratch = torch.empty(q.shape[0], self.out_features, dtype=w.dtype, device=q.device) if (w.fp4 and q.shape[0] >= 2048) else None
            gemm_b200(q, s, la, w, out=scratch, act_unsigned=self.act_unsigned, norm_q=norm_q, norm_k=norm_k, rotary_emb=rotary_emb, out_qkv=out_qkv,
                      attn_tokens=attn_tokens)
            return out_qkv
        if output is None:
            output = torch.empty(x2d.shape[0], self.out_features, dtype=x2d.dtype, device=x2d.device)
        gemm_b200(q, s, la, w, out=output, act_unsigned=self.act_unsigned, norm_q=norm_q, norm_k=norm_k, rotary_emb=rotary_emb)
        return output

    def extra_repr(self) -> str:
        return (f"in_features={self.in_features}, out_features={self.out_features}, rank={self.rank}, precision={self.precision}, "
                f"act_unsigned={self.act_unsigned}")
