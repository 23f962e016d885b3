"""CPU: the built objects really are sm_100a tcgen05 / TMEM / TMA kernels (cuobjdump -sass on nunchaku_b200/_lib/obj/*.o) -- the structural claim
DESIGN.md makes for each kernel, checked without a GPU.  SASS mnemonics (B200_PROFILING.md): tcgen05.mma = UTC*MMA, tcgen05.cp = UTCCP,
tcgen05.ld = LDTM, TMA loads / stores = UTMALDG / UTMASTG / UBLKCP, mbarriers = SYNCS; HMMA is the legacy mma.sync path."""
import collections
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "nunchaku_b200", "_lib", "obj")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not glob.glob(os.path.join(OBJ, "*.o")),
                                reason="needs cuobjdump and the objects __graft_entry__.build() leaves in nunchaku_b200/_lib/obj")


def _kernels(obj: str) -> dict:
    """mangled kernel name -> Counter of opcode prefixes"""
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in txt
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = out.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    return out


def _pick(kernels: dict, *needles: str) -> collections.Counter:
    hits = [c for name, c in kernels.items() if all(n in name for n in needles)]
    assert hits, f"no kernel matching {needles}"
    return max(hits, key=lambda c: sum(c.values()))


def test_nvfp4_cluster_gemm_is_a_tcgen05_tma_kernel():
    c = _pick(_kernels("gemm_nvfp4_cluster.o"), "gemm_nvfp4_cluster_kernel", "bfloat16")
    assert c["UTCOMMA"] >= 4          # kind::mxf4nvf4.block_scale main loop
    assert c["UTCHMMA"] >= 2          # low-rank kind::f16 MMAs into the same accumulator
    assert c["UTCCP"] >= 12           # scale factors shared memory -> TMEM
    assert c["LDTM"] >= 4 and c["UTMALDG"] >= 4 and c["SYNCS"] >= 20
    assert c["HMMA"] == 0             # no mma.sync anywhere in the GEMM


def test_single_cta_gemm_epilogues():
    k = _kernels("gemm_w4a4.o")
    fused = _pick(k, "gemm_w4a4_kernelILb1E13__nv_bfloat16Li128ELi1E")    # NVFP4, fused quantise epilogue: next layer's down projection on tcgen05
    assert fused["UTCOMMA"] >= 4 and fused["UTCHMMA"] >= 4 and fused["UTMASTG"] + fused["STG"] > 0
    litela = _pick(k, "gemm_w4a4_kernelILb1E13__nv_bfloat16Li128ELi3E")   # NVFP4, LiteLA epilogue: 8 Gram-matrix MMAs + fp32 reductions to out_vk
    assert litela["UTCHMMA"] >= 8 + 2 and litela["REDG"] >= 32
    int4 = _pick(k, "gemm_w4a4_kernelILb0E13__nv_bfloat16Li256ELi0E")     # INT4: converter warps feed kind::f16
    assert int4["UTCHMMA"] >= 4 and int4["UTCOMMA"] == 0 and int4["HMMA"] == 0


def test_attention_runs_both_gemms_on_tcgen05():
    c = _pick(_kernels("attention.o"), "attention_fp16_v2_kernel")
    assert c["UTCHMMA"] >= 16 and c["LDTM"] >= 2 and c["STTM"] >= 1 and c["UTMALDG"] >= 4   # S = Q K^T and O += P V, P written back to TMEM
    assert c["HMMA"] == 0 and c["MUFU"] >= 1


def test_quantizer_streams_by_tma_and_projects_with_mma_sync():
    c = _pick(_kernels("quantize_v2.o"), "quantize_v2_kernel", "bfloat16", "Lb1ELi16ELi28E")
    assert c["UTMALDG"] >= 1 and c["UBLKCP"] >= 1 and c["SYNCS"] >= 10 and c["HMMA"] == 32 and c["LDSM"] >= 8
