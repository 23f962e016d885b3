"""Generate golden format fixtures by running the REFERENCE's own packer (CPU torch).

Run in the authoring container only (needs /root/reference).  The reference ships no
operator-level golden vectors; the only CPU-runnable statement of this
path in the reference is `nunchaku/lora/flux/packer.py` (packed on-disk layouts of
qweight / wscales / micro-scales / bias / low-rank factors).  We import it *by file
path* (its package __init__ needs diffusers, absent here), pack seeded random tensors
and commit the inputs+outputs as small .npz fixtures.  tests/test_formats_golden.py
then pins oracle/formats.py (our closed-form restatement) against them.

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load_reference_packer():
    # stub packages so relative imports inside packer.py resolve without nunchaku/__init__.py
    for name in ("nunchaku", "nunchaku.lora", "nunchaku.lora.flux"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []  # mark as package
            sys.modules[name] = m

    def load(modname, relpath):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    load("nunchaku.utils", "nunchaku/utils.py")
    load("nunchaku.lora.flux.utils", "nunchaku/lora/flux/utils.py")
    return load("nunchaku.lora.flux.packer", "nunchaku/lora/flux/packer.py")


def bits16(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().copy()


def main():
    packer_mod = _load_reference_packer()
    packer = packer_mod.NunchakuWeightPacker(bits=4)
    g = torch.Generator().manual_seed(20260922)
    out = {}

    # ---- qweight (INT4 two's complement; NVFP4 uses identical nibble positions) ----
    N, K = 256, 384
    w = torch.randint(-8, 8, (N, K), generator=g, dtype=torch.int32)
    out["w_n256_k384"] = w.to(torch.int8).numpy()
    out["w_n256_k384_packed"] = packer.pack_weight(w.clone()).numpy()

    # ---- INT4 group scales [N, K/64] -> packed "[K/64, N]" (bf16 bit patterns) ----
    s = (torch.rand((N, K // 64), generator=g) + 0.5).to(torch.bfloat16)
    out["ws_n256_g6_bits"] = bits16(s)
    out["ws_n256_g6_packed_bits"] = bits16(packer.pack_scale(s.clone(), group_size=64))

    # ---- per-channel vectors (bias / smooth / wcscales): pack_scale(group_size=-1) ----
    b = torch.randn((N,), generator=g).to(torch.bfloat16)
    out["vec_n256_bits"] = bits16(b)
    out["vec_n256_packed_bits"] = bits16(packer.pack_scale(b.clone().view(-1, 1), group_size=-1))

    # ---- NVFP4 micro scales [N, K/16] -> packed fp8 e4m3 "[K/16, N]" ----
    ms = (torch.rand((N, K // 16), generator=g) * 4 + 0.25).to(torch.bfloat16)
    out["wms_n256_g24_bits"] = bits16(ms)
    pm = packer.pack_scale(ms.clone(), group_size=16)
    assert pm.dtype == torch.float8_e4m3fn
    out["wms_n256_g24_packed_u8"] = pm.contiguous().view(torch.uint8).numpy().copy()

    # ---- low-rank factors ----
    R = 48
    up = torch.randn((N, R), generator=g).to(torch.bfloat16)  # proj_up logical [N, R]
    out["lup_n256_r48_bits"] = bits16(up)
    out["lup_n256_r48_packed_bits"] = bits16(packer.pack_lowrank_weight(up.clone(), down=False))
    down = torch.randn((R, K), generator=g).to(torch.bfloat16)  # proj_down logical [R, K]
    out["ldown_r48_k384_bits"] = bits16(down)
    pd = packer.pack_lowrank_weight(down.clone(), down=True)
    assert tuple(pd.shape) == (K, R)
    out["ldown_r48_k384_packed_bits"] = bits16(pd)
    # reference unpack must invert its own pack (sanity of the fixture itself)
    assert torch.equal(packer.unpack_lowrank_weight(pd, down=True), down)

    # ---- packed rotary table: pack_rotemb (nunchaku/models/embeddings.py) ----
    # the module imports diffusers at top level (absent here), so only the function's own source
    # is executed, straight from the reference file (nothing is copied into this repo)
    import ast

    src_path = os.path.join(REF, "nunchaku/models/embeddings.py")
    tree = ast.parse(open(src_path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "pack_rotemb")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src_path, "exec"), ns)
    rot = torch.randn((1, 32, 64, 1, 2), generator=g, dtype=torch.float32)   # (B, M, D/2, 1, 2) = (sin, cos)
    out["rotemb_m32"] = rot.numpy()
    out["rotemb_m32_packed"] = ns["pack_rotemb"](rot.clone()).numpy()

    path = os.path.join(HERE, "packer_formats.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
