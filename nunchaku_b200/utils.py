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
