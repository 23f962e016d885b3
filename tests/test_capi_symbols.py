"""CPU: the C-ABI library loads and exports every symbol include/nunchaku_b200.h declares
(no compute calls -- there is no GPU here)."""
import os
import re

import pytest

from nunchaku_b200._C import SYMBOLS, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "nunchaku_b200.h")


def _declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by the .so"
        assert n in SYMBOLS, f"{n} has no ctypes prototype"
    assert sorted(SYMBOLS) == names


def test_abi_version_and_argument_validation_without_gpu():
    import ctypes

    from nunchaku_b200._C import GemmArgs, QuantizeArgs

    assert lib.nb200_abi_version() == 2
    # NULL / malformed arguments are rejected on the host before any CUDA call
    assert lib.nb200_quantize_w4a4_act_fuse_lora(None, None) == -1
    assert b"NULL" in lib.nb200_last_error()
    q = QuantizeArgs()
    assert lib.nb200_quantize_w4a4_act_fuse_lora(ctypes.byref(q), None) == -1
    g = GemmArgs()
    assert lib.nb200_gemm_w4a4(ctypes.byref(g), None) == -1
    assert lib.nb200_repack_qweight(None, None, 128, 128, 0, None) == -1
    assert lib.nb200_repack_qweight(1, 1, 100, 128, 0, None) == -1   # N not a multiple of 128


def test_struct_layout_matches_header_sizes():
    """ctypes mirrors of nb200_quantize_args / nb200_gemm_args: field order is taken from the
    header; sizes are what a C compiler produces for it (6*8 + 7*4 + pad + 8 + 8 -> 96; 17*8 + 11*4 + 64*4 + 2*4 -> 448)."""
    import ctypes

    from nunchaku_b200._C import GemmArgs, QuantizeArgs

    assert ctypes.sizeof(QuantizeArgs) == 104
    assert ctypes.sizeof(GemmArgs) == 544


def test_struct_layout_matches_a_c_compiler(tmp_path):
    """Compile the header with gcc and compare sizeof / offsetof of every struct field with the ctypes mirrors."""
    import ctypes
    import shutil
    import subprocess

    from nunchaku_b200._C import GemmArgs, QuantizeArgs

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    fields = {"nb200_quantize_args": QuantizeArgs, "nb200_gemm_args": GemmArgs}
    prog = ['#include "nunchaku_b200.h"', "#include <stdio.h>", "#include <stddef.h>", "int main(void) {"]
    for sname, cls in fields.items():
        prog.append(f'  printf("{sname} %zu\\n", sizeof({sname}));')
        for fname, _ in cls._fields_:
            prog.append(f'  printf("{sname}.{fname} %zu\\n", offsetof({sname}, {fname}));')
    prog += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for sname, cls in fields.items():
        assert int(out[sname]) == ctypes.sizeof(cls), sname
        for fname, _ in cls._fields_:
            assert int(out[f"{sname}.{fname}"]) == getattr(cls, fname).offset, f"{sname}.{fname}"


def test_glue_and_litela_entry_points_validate_arguments_without_a_gpu():
    """Precondition failures are reported before any CUDA call: status -1 and a message, no crash."""
    from nunchaku_b200._C import lib

    assert lib.nb200_add(1, None, None, None, 16, None) == -1
    assert b"null" in lib.nb200_last_error().lower()
    assert lib.nb200_activation(0, 1, 16, 16, 8, None) == -1            # kind must be SILU / GELU
    assert lib.nb200_activation(1, 7, 16, 16, 8, None) == -1            # unknown dtype
    assert lib.nb200_layernorm(1, 16, None, None, 16, 4, 12, 1e-6, None) == -1   # hidden % 8 != 0
    assert lib.nb200_rms_norm(1, 16, None, 16, 4, 64, 1e-6, None) == -1        # weight required
    assert lib.nb200_mul_add_batch(1, 16, None, 16, 0.0, 1, 12, 1, 12, 0, 0, 0, None) == -1   # sizes % 8
    assert lib.nb200_cast(1, 16, 1, 16, -1, None) == -1
    assert lib.nb200_litela_vk(1, 16, 16, 16, 1, 256, 100, None) == -1   # N must be 3 * heads * 32
    assert lib.nb200_linearattn_vk_mul_q(2, 16, 16, 1, 256, 4, 1e-6, None) == -1   # fp32 q is not supported
    import ctypes

    arr = (ctypes.c_void_p * 7)(*([16] * 7))
    assert lib.nb200_split_mod(1, 16, arr, 7, 70, None) == -1            # 2..6 outputs
    assert lib.nb200_add(1, 16, 16, 16, 0, None) == 0                     # empty tensors are a no-op


def test_litela_epilogue_arguments_are_validated_without_a_gpu():
    """nb200_gemm_args.out_vk (EpilogueLiteLA inside the GEMM, launch_impl:311-346): the preconditions the reference asserts (numBlocksN % 3 == 0,
    tokens a multiple of the tile height, out = relu(Q) of exactly [Mp, N/3]) come back as status -1 + message before any CUDA call."""
    import ctypes

    from nunchaku_b200._C import GemmArgs

    def args(**kw):
        g = GemmArgs()
        g.act = g.wgt = g.ascales = g.wscales = 256          # fake, 16-byte aligned: validation never dereferences
        g.Mp, g.N, g.K = 512, 384, 256
        g.dtype, g.fp4 = 1, 1
        g.out, g.M_out, g.N_out = 256, 512, 128
        g.out_vk, g.vk_tokens = 256, 256
        for k, v in kw.items():
            setattr(g, k, v)
        return g

    def rejected(g, needle):
        assert lib.nb200_gemm_w4a4(ctypes.byref(g), None) == -1
        assert needle in lib.nb200_last_error(), lib.nb200_last_error()

    rejected(args(N=256, N_out=128), b"multiple of 128")               # N / 3 not a whole number of 128-wide tiles
    rejected(args(vk_tokens=192), b"vk_tokens")                         # a 128-row tile would straddle two images
    rejected(args(vk_tokens=384), b"vk_tokens")                         # does not divide Mp
    rejected(args(N_out=384), b"[Mp, N / 3]")                           # out must be relu(Q) only
    rejected(args(out=None), b"must be non-NULL")
    rejected(args(qout=256, oscales=256, smooth_next=256), b"out_vk goes with out")
    rejected(args(rotary_emb=256, norm_q=256, norm_k=256), b"out_vk goes with out")
