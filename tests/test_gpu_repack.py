"""GPU: nb200_repack_* kernels against the closed-form formats (bit-exact)."""
import pytest
import torch

import b200_layouts as L
from oracle import formats as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rp():
    from nunchaku_b200 import repack
    from nunchaku_b200._C import check, lib

    check(lib.nb200_check_device(), "check_device")
    return repack


def test_qweight_int4_and_fp4(rp):
    g = torch.Generator().manual_seed(0)
    N, K = 256, 384
    w = torch.randint(-8, 8, (N, K), generator=g, dtype=torch.int8)
    packed = F.pack_qweight(w).cuda()
    got = rp.qweight(packed, fp4=False).cpu()
    assert torch.equal(got, L.pack_int4(w, signed=True))
    codes = (w.to(torch.int16) & 0xF).to(torch.int8)
    packed4 = F.pack_qweight(codes).cuda()
    got4 = rp.qweight(packed4, fp4=True).cpu()
    assert torch.equal(got4, L.pack_fp4(codes))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_scales_and_vectors(rp, dtype):
    g = torch.Generator().manual_seed(1)
    N, K = 384, 512
    s = (torch.rand(N, K // 64, generator=g) + 0.5).to(dtype)
    got = rp.wscales(F.pack_group_scales(s).cuda(), N, K, fp4=False).cpu()
    assert torch.equal(got, s.t().contiguous())
    ms = torch.randint(1, 120, (N, K // 16), generator=g, dtype=torch.uint8)
    gotm = rp.wscales(F.pack_micro_scales(ms).view(torch.float8_e4m3fn).cuda(), N, K, fp4=True).cpu()
    assert torch.equal(gotm, L.pack_sf_tiles(ms))
    v = torch.randn(N, generator=g).to(dtype)
    pv = F.pack_channel_vector(v).cuda()
    assert torch.equal(rp.channel_vector(pv, out_f32=False).cpu(), v)
    assert torch.equal(rp.channel_vector(pv, out_f32=True, mul=0.5).cpu(), v.float() * 0.5)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("R", [16, 32, 48])
def test_lowrank(rp, dtype, R):
    g = torch.Generator().manual_seed(2)
    N, K = 256, 384
    lu = torch.randn(N, R, generator=g).to(dtype)
    ld = torch.randn(R, K, generator=g).to(dtype)
    cs = (torch.rand(N, generator=g) + 0.5)
    got_up = rp.lora_up(F.pack_lowrank(lu, down=False).cuda(), None).cpu()
    assert torch.equal(got_up, L.lora_up_blocks(lu))
    got_up2 = rp.lora_up(F.pack_lowrank(lu, down=False).cuda(), cs.cuda()).cpu()
    exp2 = L.lora_up_blocks(lu, cs)
    # fp32 division on device vs torch: allow 1 ulp of hT
    assert torch.allclose(got_up2.float(), exp2.float(), rtol=2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10, atol=0)
    got_dn = rp.lora_down(F.pack_lowrank(ld, down=True).cuda()).cpu()
    assert torch.equal(got_dn, L.lora_down_frags(ld))
