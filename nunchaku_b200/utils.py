"""Host helpers mirroring nunchaku/utils.py for the SVDQuant path on B200."""
from __future__ import annotations

import functools

import torch


def ceil_divide(x: int, divisor: int) -> int:
    """nunchaku/utils.py:113-129."""
    return (x + divisor - 1) // divisor


def get_gpu_arch(device: str | torch.device = "cuda") -> int:
    cap = torch.cuda.get_device_capability(device)
    return cap[0] * 10 + cap[1]


def get_precision(precision: str = "auto", device: str | torch.device = "cuda", pretrained_model_name_or_path=None) -> str:
    """nunchaku/utils.py:190-231 answers "int4" for every arch except sm_120/121.  On B200 both
    paths exist here; NVFP4 is the tensor-core-native one, so "auto" prefers it unless the
    checkpoint name says otherwise (same file-name rule as the reference)."""
    assert precision in ("auto", "int4", "fp4", "nvfp4")
    if precision == "auto":
        precision = "nvfp4"
        if pretrained_model_name_or_path is not None and "int4" in str(pretrained_model_name_or_path):
            precision = "int4"
    return "nvfp4" if precision == "fp4" else precision


def check_hardware_compatibility(quantization_config: dict | None = None, device: str | torch.device = "cuda") -> None:
    """The reference raises for sm_100 (nunchaku/utils.py:308-318).  Here sm_100 is the ONLY
    supported architecture."""
    arch = get_gpu_arch(device)
    if arch // 10 != 10:
        raise ValueError(f"nunchaku_b200 supports only sm_100 (B200); found sm_{arch}.")


def torch_dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float16:
        return 0
    if dtype == torch.bfloat16:
        return 1
    raise TypeError(f"SVDQuant W4A4 runs in float16 or bfloat16, got {dtype}")


def on_device_of(arg_name: str, position: int = 0):
    """Decorator: run the op with the CUDA device of the named tensor argument current, so that ``torch.cuda.current_stream()``
    and the kernel launch refer to the device that owns the pointers (the reference's ops take the device from their tensors
    the same way, src/interop/torch.cpp:84-91).  CPU / missing tensors fall through to the op's own error handling."""

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            t = kwargs.get(arg_name, args[position] if len(args) > position else None)
            if isinstance(t, torch.Tensor) and t.is_cuda:
                with torch.cuda.device(t.device):
                    return fn(*args, **kwargs)
            return fn(*args, **kwargs)

        return wrapper

    return deco


def pack_rotemb(sin: torch.Tensor, cos: torch.Tensor) -> torch.Tensor:
    """Rotary table in the layout the RMSNorm+RoPE epilogue reads (the reference's ``pack_rotemb``,
    nunchaku/models/transformers/transformer_flux.py:60-92; consumer epilogues.cuh:283-300): ``sin`` / ``cos`` are [M, 64] fp32,
    pair i rotating head columns (2i, 2i+1); M a multiple of 16.  Returns fp32 [M, 128].

    The packed order is the fp32 accumulator-fragment order of a 16-row tile: [M/16][pair/4][row%8][pair%4][(row%16)/8][sin|cos],
    which is a pure axis permutation of the logical [M/16][(row%16)/8][row%8][pair/4][pair%4][sin|cos]."""
    M, P = sin.shape
    if P != 64 or M % 16 != 0 or cos.shape != sin.shape:
        raise ValueError("pack_rotemb: sin / cos must be [M, 64] with M a multiple of 16")
    t = torch.stack([sin.to(torch.float32), cos.to(torch.float32)], dim=-1)       # [M, 64, 2]
    t = t.view(M // 16, 2, 8, 16, 4, 2).permute(0, 3, 2, 4, 1, 5)                  # -> [mb, pair/4, row%8, pair%4, hi, s]
    return t.reshape(M, 128).contiguous()


def attach(model) -> int:
    """Bind-onto-the-reference helper: after the operator functions have been patched onto the
    reference's modules, install ``load_state_dict`` hooks on every module that owns SVDQuant parameters so that the converted
    weight copies (``nunchaku_b200.repack`` cache) are dropped whenever a checkpoint / LoRA is loaded in place.  Returns the
    number of modules hooked."""
    from . import repack

    names = ("qweight", "wscales", "bias", "smooth_factor", "proj_down", "proj_up", "wcscales")
    hooked = 0
    for mod in model.modules():
        if hasattr(mod, "qweight") and hasattr(mod, "proj_up"):
            def _drop(module, _keys, _names=names):
                for n in _names:
                    p = getattr(module, n, None)
                    if isinstance(p, torch.Tensor):
                        repack.invalidate(p)
            mod.register_load_state_dict_post_hook(_drop)
            hooked += 1
    return hooked
