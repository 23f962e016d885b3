"""Pin oracle/formats.py against fixtures produced by the reference's own packer.py."""
import os

import numpy as np
import pytest
import torch

from oracle import formats as F

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "packer_formats.npz"))


def _bf16(bits):
    return torch.from_numpy(bits.copy()).view(torch.bfloat16)


def test_qweight_pack_matches_reference():
    w = torch.from_numpy(GOLD["w_n256_k384"])
    ref = torch.from_numpy(GOLD["w_n256_k384_packed"])
    assert torch.equal(F.pack_qweight(w), ref)
    assert torch.equal(F.unpack_qweight(ref, signed=True), w)
    # raw (unsigned) codes view used for NVFP4
    assert torch.equal(F.unpack_qweight(ref, signed=False).to(torch.int16), (w.to(torch.int16) & 0xF))


def test_group_scales_match_reference():
    s = _bf16(GOLD["ws_n256_g6_bits"])
    ref = _bf16(GOLD["ws_n256_g6_packed_bits"])
    assert torch.equal(F.pack_group_scales(s).view(torch.int16), ref.view(torch.int16))
    assert torch.equal(F.unpack_group_scales(ref).view(torch.int16), s.view(torch.int16))


def test_channel_vector_matches_reference():
    v = _bf16(GOLD["vec_n256_bits"])
    ref = _bf16(GOLD["vec_n256_packed_bits"])
    assert torch.equal(F.pack_channel_vector(v).view(torch.int16), ref.view(torch.int16))
    assert torch.equal(F.unpack_channel_vector(ref).view(torch.int16), v.view(torch.int16))


def test_micro_scales_match_reference():
    ms = _bf16(GOLD["wms_n256_g24_bits"]).to(torch.float8_e4m3fn)
    ref = torch.from_numpy(GOLD["wms_n256_g24_packed_u8"])
    assert torch.equal(F.pack_micro_scales(ms).view(torch.uint8), ref)
    assert torch.equal(F.unpack_micro_scales(ref), ms.view(torch.uint8))


@pytest.mark.parametrize("down", [False, True])
def test_lowrank_matches_reference(down):
    if down:
        logical = _bf16(GOLD["ldown_r48_k384_bits"])
        ref = _bf16(GOLD["ldown_r48_k384_packed_bits"])
    else:
        logical = _bf16(GOLD["lup_n256_r48_bits"])
        ref = _bf16(GOLD["lup_n256_r48_packed_bits"])
    assert torch.equal(F.pack_lowrank(logical, down).view(torch.int16), ref.view(torch.int16))
    assert torch.equal(F.unpack_lowrank(ref, down).view(torch.int16), logical.view(torch.int16))


def test_roundtrips_random_shapes():
    g = torch.Generator().manual_seed(1)
    for N, K in [(128, 128), (384, 640)]:
        w = torch.randint(-8, 8, (N, K), generator=g, dtype=torch.int8)
        assert torch.equal(F.unpack_qweight(F.pack_qweight(w)), w)
        s = torch.randn(N, K // 64, generator=g).to(torch.float16)
        assert torch.equal(F.unpack_group_scales(F.pack_group_scales(s)), s)
        m = torch.randint(0, 127, (N, K // 16), generator=g, dtype=torch.uint8)
        assert torch.equal(F.unpack_micro_scales(F.pack_micro_scales(m)), m)


def test_pack_rotemb_matches_reference():
    rot = torch.from_numpy(GOLD["rotemb_m32"])            # (1, M, 64, 1, 2) = (sin, cos)
    ref = torch.from_numpy(GOLD["rotemb_m32_packed"])     # (1, M, 128)
    sin, cos = rot[0, :, :, 0, 0], rot[0, :, :, 0, 1]
    assert torch.equal(F.pack_rotemb(sin, cos), ref[0])
