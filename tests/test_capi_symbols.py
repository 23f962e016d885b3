"""CPU: the C-ABI library loads and exports every symbol include/nunchaku_b200.h declares
(no compute calls -- there is no GPU here)."""
import os
import re

from nunchaku_b200._C import SYMBOLS, lib

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

    assert lib.nb200_abi_version() == 1
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
    assert ctypes.sizeof(GemmArgs) == 512
