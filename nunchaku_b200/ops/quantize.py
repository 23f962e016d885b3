"""Mirror of nunchaku/ops/quantize.py:10-80 (svdq_quantize_w4a4_act_fuse_lora_cuda) on B200."""
from __future__ import annotations

import ctypes

import torch

from .. import repack
from .._C import QuantizeArgs, check, lib
from ..utils import ceil_divide, on_device_of, torch_dtype_code


def _launch(input: torch.Tensor, output, oscales, lora_act_out, lora_down_b200, smooth_b200, rank: int, fuse_glu: bool, fp4: bool,
            pad_size: int, shift_unsigned: bool):
    """allocate missing outputs (nunchaku/ops/quantize.py:60-78 shapes) and call nb200_quantize_w4a4_act_fuse_lora"""
    if input.dim() != 2:
        raise ValueError("input must be 2-D (M, K)")
    if not input.is_cuda:
        raise RuntimeError("nunchaku_b200 has no CPU path: input must be a CUDA tensor")
    if pad_size % 256 != 0:
        raise ValueError("pad_size must be a multiple of 256 (kernel tile, launch_impl:462-463)")
    if shift_unsigned and (fp4 or fuse_glu):
        raise ValueError("shift_unsigned is INT4 only and excludes fuse_glu")
    input = input.contiguous()
    batch_size, channels = input.shape
    if fuse_glu:
        channels //= 2
    group = 16 if fp4 else 64
    if channels % 128 != 0:
        raise ValueError("the channel count must be a multiple of 128")
    batch_size_pad = ceil_divide(batch_size, pad_size) * pad_size
    dev = input.device
    if output is None:
        output = torch.empty(batch_size_pad, channels // 2, dtype=torch.uint8, device=dev)
    if oscales is None:
        oscales = torch.empty(channels // group, batch_size_pad, dtype=torch.float8_e4m3fn if fp4 else input.dtype, device=dev)
    if lora_act_out is None:
        lora_act_out = torch.empty(batch_size_pad, rank, dtype=torch.float32, device=dev)
    if not (output.is_contiguous() and output.shape[0] == batch_size_pad and output.shape[-1] * 2 == channels):
        raise ValueError("output must be contiguous [M_pad, K/2]")
    if not (oscales.is_contiguous() and oscales.numel() == (channels // group) * batch_size_pad):
        raise ValueError("oscales must be contiguous [K/G, M_pad]")
    if not (lora_act_out.is_contiguous() and tuple(lora_act_out.shape) == (batch_size_pad, rank) and lora_act_out.dtype == torch.float32):
        raise ValueError("lora_act_out must be contiguous float32 [M_pad, rank]")
    args = QuantizeArgs()
    args.input, args.output, args.oscales = input.data_ptr(), output.data_ptr(), oscales.data_ptr()
    args.lora_down = None if rank == 0 else lora_down_b200.data_ptr()
    args.lora_act_out = None if rank == 0 else lora_act_out.data_ptr()
    args.smooth = None if smooth_b200 is None else smooth_b200.data_ptr()
    args.M, args.Mp, args.K, args.R = batch_size, batch_size_pad, channels, rank
    args.dtype = torch_dtype_code(input.dtype)
    args.fuse_glu, args.fp4 = int(fuse_glu), int(fp4)
    args.act_unsigned_shift = int(shift_unsigned)
    ws = _workspace(batch_size_pad, channels, dev)
    args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
    check(lib.nb200_quantize_w4a4_act_fuse_lora(ctypes.byref(args), torch.cuda.current_stream().cuda_stream), "quantize_w4a4_act_fuse_lora")
    return output, oscales, lora_act_out


@on_device_of("input")
def quantize_b200(input: torch.Tensor, w, *, output=None, oscales=None, lora_act_out=None, fuse_glu: bool = False, pad_size: int = 256,
                  shift_unsigned: bool = False):
    """Quantize for the converted layer ``w`` (``nunchaku_b200.weights.B200Weights``): its smoothing vector, its low-rank down factor."""
    return _launch(input, output, oscales, lora_act_out, w.lora_down, w.smooth, w.rank, fuse_glu, w.fp4, pad_size, shift_unsigned)


@on_device_of("input")
def svdq_quantize_w4a4_act_fuse_lora_cuda(
    input: torch.Tensor,
    output: torch.Tensor | None = None,
    oscales: torch.Tensor | None = None,
    lora_down: torch.Tensor | None = None,
    lora_act_out: torch.Tensor | None = None,
    smooth: torch.Tensor | None = None,
    fuse_glu: bool = False,
    fp4: bool = False,
    pad_size: int = 256,
    *,
    shift_unsigned: bool = False,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Quantize activations to 4 bit and compute the low-rank down projection; same arguments, shapes, dtypes and return value
    as the reference wrapper (nunchaku/ops/quantize.py:10-80).  ``lora_down`` / ``smooth`` arrive in the reference's packed
    checkpoint layout (converted through ``nunchaku_b200.repack``); the three outputs are only meaningful to this package's GEMM
    (their internal element order is the B200 inter-op layout, exactly as the reference's outputs are in its own fragment order)::

        output        (M_pad, K // 2)   uint8
        oscales       (K // G, M_pad)   input dtype (INT4, G=64) | float8_e4m3fn (NVFP4, G=16)
        lora_act_out  (M_pad, R)        float32

    ``shift_unsigned`` (keyword-only, INT4, not in the reference signature): quantise ``(x + 0.171875) / smooth`` to unsigned codes
    with scale = max / 15 -- what the reference's fused GELU epilogue produces for fc2 (launch_impl:282-310) -- while the low-rank
    projection still sees ``x``; the consumer GEMM must be called with ``act_unsigned=True``.
    """
    if lora_down is None:
        raise ValueError("lora_down is required (the reference dereferences it unconditionally)")
    if input.dim() != 2:
        raise ValueError("input must be 2-D (M, K)")
    if not input.is_cuda:
        raise RuntimeError("nunchaku_b200 has no CPU path: input must be a CUDA tensor")
    rank = lora_down.shape[1]
    return _launch(input, output, oscales, lora_act_out, repack.lora_down(lora_down) if rank > 0 else None,
                   None if smooth is None else repack.channel_vector(smooth, out_f32=False), rank, fuse_glu, fp4, pad_size, shift_unsigned)


_ws_cache: dict = {}


def _workspace(Mp: int, K: int, device) -> torch.Tensor:
    """Zero-initialised scratch for the small-M split-K path (kernel leaves it clean); one per
    (device, stream) and grown on demand -- allocate before CUDA-graph capture by warming up."""
    need = int(lib.nb200_quantize_workspace_bytes(Mp, K))
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws
