"""One-time repack of reference-layout parameters into the B200 layouts (device side, via the
C-ABI ``nb200_repack_*`` kernels) with a per-parameter cache.

The reference keeps checkpoints in mma.sync fragment order (nunchaku/lora/flux/packer.py);
its loader hands those tensors straight to the kernels (src/Linear.cpp:124-154).  Here the
first use of a parameter converts it once; the cache is keyed on the tensor's storage address,
shape, dtype and ``_version`` so in-place updates (``load_state_dict``, LoRA hot-swap, which
re-allocates lora_up/down with a new rank -- Linear.cpp:124-134) invalidate the entry.
"""
from __future__ import annotations

import weakref

import torch

from ._C import check, lib
from .utils import torch_dtype_code

_cache: dict[tuple, tuple] = {}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _key(kind: str, t: torch.Tensor) -> tuple:
    return (kind, t.device.index, t.data_ptr())


def _sig(t: torch.Tensor, *extra) -> tuple:
    return (tuple(t.shape), t.dtype, t._version, *extra)


def _lookup(kind: str, t: torch.Tensor, sig: tuple):
    hit = _cache.get(_key(kind, t))
    if hit is not None and hit[0] == sig:
        return hit[1]
    return None


def _store(kind: str, t: torch.Tensor, sig: tuple, value):
    key = _key(kind, t)
    _cache[key] = (sig, value)
    try:  # drop the entry when the source tensor object dies (its address may be reused)
        weakref.finalize(t, _cache.pop, key, None)
    except TypeError:  # pragma: no cover
        pass
    return value


def clear_cache() -> None:
    _cache.clear()


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (nunchaku_b200 has no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def qweight(wgt: torch.Tensor, fp4: bool) -> torch.Tensor:
    """reference int8 [N, K/2] -> B200 u8 [N, K/2]."""
    _require_cuda(wgt, "wgt")
    sig = _sig(wgt, fp4)
    hit = _lookup("qweight", wgt, sig)
    if hit is not None:
        return hit
    N, Kh = wgt.shape
    out = torch.empty(N, Kh, dtype=torch.uint8, device=wgt.device)
    check(lib.nb200_repack_qweight(wgt.data_ptr(), out.data_ptr(), N, Kh * 2, int(fp4), _stream()), "repack_qweight")
    return _store("qweight", wgt, sig, out)


def wscales(ws: torch.Tensor, N: int, K: int, fp4: bool) -> torch.Tensor:
    _require_cuda(ws, "wscales")
    sig = _sig(ws, fp4, N, K)
    hit = _lookup("wscales", ws, sig)
    if hit is not None:
        return hit
    if fp4:
        assert ws.numel() == N * K // 16 and ws.element_size() == 1, "NVFP4 wscales must be [K/16, N] fp8"
        out = torch.empty(N * K // 16, dtype=torch.uint8, device=ws.device)
        check(lib.nb200_repack_wscales_fp4(ws.data_ptr(), out.data_ptr(), N, K, _stream()), "repack_wscales_fp4")
    else:
        assert ws.numel() == N * K // 64 and ws.element_size() == 2, "INT4 wscales must be [K/64, N] fp16/bf16"
        out = torch.empty(K // 64, N, dtype=ws.dtype, device=ws.device)
        check(lib.nb200_repack_wscales_int4(ws.data_ptr(), out.data_ptr(), N, K, _stream()), "repack_wscales_int4")
    return _store("wscales", ws, sig, out)


def channel_vector(v: torch.Tensor, out_f32: bool, mul: float = 1.0) -> torch.Tensor:
    """bias / smooth_factor / wcscales (pack_scale(group_size=-1) order) -> natural order."""
    _require_cuda(v, "channel vector")
    sig = _sig(v, out_f32, float(mul))
    hit = _lookup("vec", v, sig)
    if hit is not None:
        return hit
    N = v.numel()
    out = torch.empty(N, dtype=torch.float32 if out_f32 else v.dtype, device=v.device)
    check(
        lib.nb200_repack_channel_vector(v.data_ptr(), out.data_ptr(), N, torch_dtype_code(v.dtype), int(out_f32),
                                        float(mul), _stream()),
        "repack_channel_vector",
    )
    return _store("vec", v, sig, out)


def lora_up(lu: torch.Tensor, cscale: torch.Tensor | None) -> torch.Tensor:
    """reference [N, R] -> UMMA K-major blocks, divided by cscale[n] (alpha * wcscales)."""
    _require_cuda(lu, "lora_up")
    sig = _sig(lu, None if cscale is None else (cscale.data_ptr(), cscale._version))
    hit = _lookup("lora_up", lu, sig)
    if hit is not None:
        return hit
    N, R = lu.shape
    Rp = (R + 31) // 32 * 32
    out = torch.empty(N * Rp, dtype=lu.dtype, device=lu.device)
    check(
        lib.nb200_repack_lora_up(lu.data_ptr(), out.data_ptr(), None if cscale is None else cscale.data_ptr(), N, R,
                                 torch_dtype_code(lu.dtype), _stream()),
        "repack_lora_up",
    )
    return _store("lora_up", lu, sig, out)


def lora_down(ld: torch.Tensor) -> torch.Tensor:
    """reference [K, R] -> mma.sync B-fragment order of the quantize kernel."""
    _require_cuda(ld, "lora_down")
    sig = _sig(ld)
    hit = _lookup("lora_down", ld, sig)
    if hit is not None:
        return hit
    K, R = ld.shape
    Rp = (R + 31) // 32 * 32
    out = torch.empty(2 * K * Rp, dtype=ld.dtype, device=ld.device)
    check(lib.nb200_repack_lora_down(ld.data_ptr(), out.data_ptr(), K, R, torch_dtype_code(ld.dtype), _stream()),
          "repack_lora_down")
    return _store("lora_down", ld, sig, out)


def lora_down_next(ld: torch.Tensor) -> torch.Tensor:
    """reference [K, R] -> logical [R, K] row-major: TMA source of the fused fc1->fc2 down projection."""
    _require_cuda(ld, "lora_down")
    sig = _sig(ld)
    hit = _lookup("lora_down_next", ld, sig)
    if hit is not None:
        return hit
    K, R = ld.shape
    out = torch.empty(R, K, dtype=ld.dtype, device=ld.device)
    check(lib.nb200_repack_lora_down_next(ld.data_ptr(), out.data_ptr(), K, R, torch_dtype_code(ld.dtype), _stream()),
          "repack_lora_down_next")
    return _store("lora_down_next", ld, sig, out)
