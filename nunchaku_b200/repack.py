"""One-time conversion of reference-layout parameters into the B200 layouts (device side, via the C-ABI
``nb200_repack_*`` kernels).

The reference keeps checkpoints in mma.sync fragment order (nunchaku/lora/flux/packer.py) and its loader hands those
tensors straight to the kernels (src/Linear.cpp:124-154).  Two ways to get the converted copies:

* ``nunchaku_b200.weights.B200Weights.from_reference`` (used by ``SVDQW4A4Linear``): ``cache=False`` conversions owned by
  the module, rebuilt when the module is told its parameters changed (``load_state_dict`` hook / ``invalidate()``).
* the raw-tensor operator calls (``svdq_gemm_w4a4_cuda(act, wgt, ...)`` -- the reference's own signature, used when the ops
  are bound onto the reference's classes): a cache keyed on the source tensor's address, shape, dtype and ``_version``.
  ``_version`` does NOT see ``param.data.copy_()`` / ``param.data.mul_()`` (what checkpoint loaders and LoRA updates
  do), so a loader must call ``invalidate(param)`` (or ``clear_cache()``) after touching parameters in place;
  ``nunchaku_b200.utils.attach(model)`` installs ``load_state_dict`` hooks that do.  Entries are also dropped when the
  source tensor object dies, and when the same tensor object shows up at a new address (``param.data = new``).
"""
from __future__ import annotations

import weakref

import torch

from ._C import check, lib
from .utils import torch_dtype_code

_cache: dict[tuple, tuple] = {}      # (kind, device, data_ptr) -> (signature, converted tensor, id(source))
_by_source: dict[int, set] = {}      # id(source tensor) -> keys it produced


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _key(kind: str, t: torch.Tensor) -> tuple:
    return (kind, t.device.index, t.data_ptr())


def _version(t: torch.Tensor) -> int:
    try:
        return t._version
    except RuntimeError:   # inference-mode tensors do not track versions
        return -1


def _sig(t: torch.Tensor, *extra) -> tuple:
    return (tuple(t.shape), t.dtype, _version(t), *extra)


def _lookup(kind: str, t: torch.Tensor, sig: tuple):
    hit = _cache.get(_key(kind, t))
    if hit is not None and hit[0] == sig:
        return hit[1]
    return None


def _drop_keys(keys) -> None:
    for k in list(keys):
        _cache.pop(k, None)


def _store(kind: str, t: torch.Tensor, sig: tuple, value):
    key = _key(kind, t)
    sid = id(t)
    mine = _by_source.get(sid)
    if mine is None:
        mine = _by_source[sid] = set()
        try:   # drop every entry of this source when the tensor object dies (its address may be reused)
            weakref.finalize(t, lambda s=sid: _drop_keys(_by_source.pop(s, ())))
        except TypeError:  # pragma: no cover
            pass
    # the same tensor object at a NEW address (param.data = other): its old entries are stale and pin GPU memory
    stale = {k for k in mine if k[0] == kind and k != key}
    _drop_keys(stale)
    mine -= stale
    mine.add(key)
    _cache[key] = (sig, value, sid)
    return value


def invalidate(t: torch.Tensor) -> None:
    """Forget every converted copy of ``t`` (call after changing a parameter in place: ``.data.copy_``, LoRA merge ...)."""
    ptr = t.data_ptr()
    dev = t.device.index
    for k in [k for k in _cache if k[1] == dev and k[2] == ptr]:
        _cache.pop(k, None)
    _drop_keys(_by_source.pop(id(t), ()))


def clear_cache() -> None:
    _cache.clear()
    _by_source.clear()


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (nunchaku_b200 has no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _cached(kind: str, t: torch.Tensor, sig: tuple, cache: bool, make):
    if cache:
        hit = _lookup(kind, t, sig)
        if hit is not None:
            return hit
    with torch.cuda.device(t.device):
        value = make()
    return _store(kind, t, sig, value) if cache else value


def qweight(wgt: torch.Tensor, fp4: bool, cache: bool = True) -> torch.Tensor:
    """reference int8 [N, K/2] -> B200 u8 [N, K/2]."""
    _require_cuda(wgt, "wgt")
    N, Kh = wgt.shape

    def make():
        out = torch.empty(N, Kh, dtype=torch.uint8, device=wgt.device)
        check(lib.nb200_repack_qweight(wgt.data_ptr(), out.data_ptr(), N, Kh * 2, int(fp4), _stream()), "repack_qweight")
        return out

    return _cached("qweight", wgt, _sig(wgt, fp4), cache, make)


def wscales(ws: torch.Tensor, N: int, K: int, fp4: bool, cache: bool = True) -> torch.Tensor:
    _require_cuda(ws, "wscales")

    def make():
        if fp4:
            if ws.numel() != N * K // 16 or ws.element_size() != 1:
                raise ValueError("NVFP4 wscales must be [K/16, N] fp8")
            out = torch.empty(N * K // 16, dtype=torch.uint8, device=ws.device)
            check(lib.nb200_repack_wscales_fp4(ws.data_ptr(), out.data_ptr(), N, K, _stream()), "repack_wscales_fp4")
        else:
            if ws.numel() != N * K // 64 or ws.element_size() != 2:
                raise ValueError("INT4 wscales must be [K/64, N] fp16/bf16")
            out = torch.empty(K // 64, N, dtype=ws.dtype, device=ws.device)
            check(lib.nb200_repack_wscales_int4(ws.data_ptr(), out.data_ptr(), N, K, _stream()), "repack_wscales_int4")
        return out

    return _cached("wscales", ws, _sig(ws, fp4, N, K), cache, make)


def channel_vector(v: torch.Tensor, out_f32: bool, mul: float = 1.0, cache: bool = True) -> torch.Tensor:
    """bias / smooth_factor / wcscales (pack_scale(group_size=-1) order) -> natural order."""
    _require_cuda(v, "channel vector")
    N = v.numel()

    def make():
        out = torch.empty(N, dtype=torch.float32 if out_f32 else v.dtype, device=v.device)
        check(lib.nb200_repack_channel_vector(v.data_ptr(), out.data_ptr(), N, torch_dtype_code(v.dtype), int(out_f32), float(mul), _stream()),
              "repack_channel_vector")
        return out

    return _cached("vec", v, _sig(v, out_f32, float(mul)), cache, make)


def lora_up(lu: torch.Tensor, cscale: torch.Tensor | None, cache: bool = True) -> torch.Tensor:
    """reference [N, R] -> UMMA K-major blocks, divided by cscale[n] (alpha * wcscales)."""
    _require_cuda(lu, "lora_up")
    N, R = lu.shape

    def make():
        out = torch.empty(N * ((R + 31) // 32 * 32), dtype=lu.dtype, device=lu.device)
        check(lib.nb200_repack_lora_up(lu.data_ptr(), out.data_ptr(), None if cscale is None else cscale.data_ptr(), N, R,
                                       torch_dtype_code(lu.dtype), _stream()), "repack_lora_up")
        return out

    return _cached("lora_up", lu, _sig(lu, None if cscale is None else (cscale.data_ptr(), _version(cscale))), cache, make)


def lora_down(ld: torch.Tensor, cache: bool = True) -> torch.Tensor:
    """reference [K, R] -> mma.sync B-fragment order of the quantize kernel."""
    _require_cuda(ld, "lora_down")
    K, R = ld.shape

    def make():
        out = torch.empty(2 * K * ((R + 31) // 32 * 32), dtype=ld.dtype, device=ld.device)
        check(lib.nb200_repack_lora_down(ld.data_ptr(), out.data_ptr(), K, R, torch_dtype_code(ld.dtype), _stream()), "repack_lora_down")
        return out

    return _cached("lora_down", ld, _sig(ld), cache, make)


def lora_down_next(ld: torch.Tensor, cache: bool = True) -> torch.Tensor:
    """reference [K, R] -> logical [R, K] row-major: TMA source of the fused fc1->fc2 down projection."""
    _require_cuda(ld, "lora_down")
    K, R = ld.shape

    def make():
        out = torch.empty(R, K, dtype=ld.dtype, device=ld.device)
        check(lib.nb200_repack_lora_down_next(ld.data_ptr(), out.data_ptr(), K, R, torch_dtype_code(ld.dtype), _stream()),
              "repack_lora_down_next")
        return out

    return _cached("lora_down_next", ld, _sig(ld), cache, make)
