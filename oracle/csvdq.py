"""TEST INFRASTRUCTURE (oracle) -- ctypes loader of the plain-C restatement oracle/svdq_ref.c.

`build()` compiles it with gcc (-O2 -fopenmp) into oracle/_build/libsvdq_ref.so; `linear_forward(layer, x)` mirrors
oracle.svdq.svdq_linear_forward(layer, x, mode="ref").  Only tests/, __graft_entry__ and bench.py may import this module."""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

import numpy as np
import torch

from . import svdq as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "svdq_ref.c")
LIB = os.path.join(HERE, "_build", "libsvdq_ref.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: the C oracle cannot be built")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run([gcc, "-O2", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        fp, i8 = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int8)
        _lib.svdq_linear_ref.restype = None
        _lib.svdq_linear_ref.argtypes = [ctypes.c_int] * 6 + [fp, fp, fp, i8, fp, fp, fp, ctypes.c_float, fp, fp]
        _lib.svdq_set_threads.restype = None
        _lib.svdq_set_threads.argtypes = [ctypes.c_int]
    return _lib


def set_threads(n: int) -> None:
    """OpenMP team size of the C oracle (torchrun exports OMP_NUM_THREADS=1)."""
    _load().svdq_set_threads(int(n))


def _f32(t: torch.Tensor | None):
    if t is None:
        return None, None
    a = np.ascontiguousarray(t.detach().to(torch.float32).cpu().numpy())
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def linear_forward(layer: "O.SynthLayer", x: torch.Tensor) -> torch.Tensor:
    """Reference-emulating forward of one SVDQuant linear in C (all host threads via OpenMP)."""
    lib = _load()
    hT = layer.hT
    M, K = x.shape
    N, R = layer.qw.shape[0], layer.lora_up.shape[1]
    ws = O.e4m3_decode(layer.wscales).to(torch.float32) if layer.fp4 else layer.wscales
    keep = []
    ptr = {}
    for name, t in (("x", x), ("smooth", layer.smooth), ("ld", layer.lora_down), ("ws", ws), ("bias", layer.bias),
                    ("wcs", layer.wcscales), ("lu", layer.lora_up)):
        a, p = _f32(t)
        keep.append(a)
        ptr[name] = p
    qw = np.ascontiguousarray(layer.qw.to(torch.int8).numpy())
    out = np.empty((M, N), dtype=np.float32)
    lib.svdq_linear_ref(M, K, N, R, int(layer.fp4), int(hT == torch.bfloat16), ptr["x"], ptr["smooth"], ptr["ld"],
                        qw.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)), ptr["ws"], ptr["bias"], ptr["wcs"], float(layer.alpha), ptr["lu"],
                        out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return torch.from_numpy(out).to(hT)
