"""GPU parity: quantize_w4a4_act_fuse_lora (CUDA, through the C ABI) vs the CPU oracle.

Tolerances (written here, per the contract):
  * scales: bit-exact (they depend only on the exact group absmax);
  * 4-bit codes: the kernel uses div.approx / rcp.approx.ftz (<= 2 ulp) where the oracle divides
    exactly, so an element sitting on a rounding boundary may move one step:
    mismatch fraction <= 2e-3 and |step| <= 1 (SURVEY.md A.5 note); with smooth=None there is
    no division before rounding except the reciprocal, same rule;
  * lora_act: fp32 accumulation order differs: rel-Frobenius <= 1e-5; run-to-run bit-identical.
"""
import pytest
import torch

import b200_layouts as L
from gpu_util import diag
from oracle import formats as F
from oracle import svdq as O

pytestmark = pytest.mark.gpu


def _run(x, smooth, ld, fp4, fuse_glu=False):
    from nunchaku_b200.ops.quantize import svdq_quantize_w4a4_act_fuse_lora_cuda as qop

    sm = None if smooth is None else F.pack_channel_vector(smooth).cuda()
    q, s, la = qop(x.cuda(), lora_down=F.pack_lowrank(ld, down=True).cuda(), smooth=sm, fp4=fp4, fuse_glu=fuse_glu)
    torch.cuda.synchronize()
    return q, s, la


def _check(x, smooth, ld, fp4, fuse_glu=False):
    ref = O.quantize_w4a4_act_fuse_lora(x, smooth, ld, fp4=fp4, fuse_glu=fuse_glu)
    q, s, la = _run(x, smooth, ld, fp4, fuse_glu)
    Mp, K = ref.q.shape
    assert q.shape == (Mp, K // 2) and q.dtype == torch.uint8
    if fp4:
        codes = L.unpack_fp4(q.cpu())
        scales = L.unpack_sf_tiles(s.cpu().view(torch.uint8).reshape(-1), Mp, K // 16).t().contiguous()
        assert s.shape == (K // 16, Mp) and s.dtype == torch.float8_e4m3fn
        assert torch.equal(scales, ref.scales), diag("fp4 scales", scales.float(), ref.scales.float())
        # compare dequantised values where scale == 0 (codes are don't-care there)
        nz = (O.e4m3_decode(ref.scales).t() != 0).repeat_interleave(16, dim=1)
        cmp = O.compare_codes(codes[nz], ref.q[nz], fp4=True)
    else:
        codes = L.unpack_int4(q.cpu(), signed=True)
        assert s.shape == (K // 64, Mp) and s.dtype == x.dtype
        assert torch.equal(s.cpu().view(torch.int16), ref.scales.view(torch.int16)), diag("int4 scales", s.float(), ref.scales.float())
        cmp = O.compare_codes(codes, ref.q, fp4=False)
    assert cmp["frac"] <= 2e-3 and cmp["max_step"] <= 1, (cmp, diag("codes", codes.float(), ref.q.float()))
    e = O.rel_fro(la.cpu(), ref.lora_act)
    assert e <= 1e-5, diag("lora_act", la, ref.lora_act)
    assert torch.all(la[ref.M:] == 0)
    # determinism
    q2, s2, la2 = _run(x, smooth, ld, fp4, fuse_glu)
    assert torch.equal(q, q2) and torch.equal(la, la2) and torch.equal(s.view(torch.uint8), s2.view(torch.uint8))
    return cmp


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,K,R", [(1, 128, 16), (300, 384, 32), (256, 3072, 32), (777, 1024, 48)])
def test_quantize_matches_oracle(fp4, hT, M, K, R):
    g = torch.Generator().manual_seed(M * 7 + K)
    smooth = torch.exp(torch.randn(K, generator=g) * 0.5).clamp(0.1, 10).to(hT)
    x = O.make_activations(M, K, hT, seed=M + K, smooth=smooth)
    ld = (torch.randn(R, K, generator=g) * 0.05).to(hT)
    _check(x, smooth, ld, fp4)


@pytest.mark.parametrize("fp4", [False, True])
def test_quantize_edge_cases(fp4):
    hT = torch.bfloat16
    K, R = 256, 16
    g = torch.Generator().manual_seed(3)
    ld = (torch.randn(R, K, generator=g) * 0.05).to(hT)
    # all-zero rows / groups, exact ties, big outliers, no smoothing
    x = torch.zeros(40, K, dtype=hT)
    x[1, :64] = torch.linspace(-7, 7, 64).to(hT)
    x[2, 64:128] = 3.5
    x[2, 64] = 7.0
    x[3] = (torch.randn(K, generator=g) * 1000).to(hT)
    x[4, 5] = 60000.0 if False else 3.0e4
    _check(x, None, ld, fp4)


@pytest.mark.parametrize("fp4", [False, True])
def test_quantize_fuse_glu(fp4):
    hT = torch.bfloat16
    M, K, R = 130, 256, 16
    g = torch.Generator().manual_seed(4)
    x = torch.randn(M, 2 * K, generator=g).to(hT)
    smooth = (torch.rand(K, generator=g) + 0.5).to(hT)
    ld = (torch.randn(R, K, generator=g) * 0.05).to(hT)
    ref = O.quantize_w4a4_act_fuse_lora(x, smooth, ld, fp4=fp4, fuse_glu=True)
    q, s, la = _run(x, smooth, ld, fp4, fuse_glu=True)
    # silu uses ex2.approx/rcp.approx on device: GLU outputs may differ by an hT ulp, which moves
    # group maxima -> compare dequantised tensors norm-wise instead of codes
    if fp4:
        codes = L.unpack_fp4(q.cpu())
        sc = L.unpack_sf_tiles(s.cpu().view(torch.uint8).reshape(-1), ref.q.shape[0], K // 16)
        deq = O.dequant(codes, sc, True)
    else:
        deq = O.dequant(L.unpack_int4(q.cpu(), True), s.cpu().t().contiguous(), False)
    deq_ref = O.dequant(ref.q, ref.scales.t().contiguous(), fp4)
    assert O.rel_fro(deq, deq_ref) < 2e-2, diag("glu deq", deq, deq_ref)
    assert O.rel_fro(la.cpu(), ref.lora_act) < 5e-3, diag("glu lora", la, ref.lora_act)


@pytest.mark.parametrize("fp4", [False, True])
def test_quantize_rank_zero(fp4):
    """the reference's GEMM_W4A4 starts with lora_rank = 0 (src/Linear.cpp:92-117) and only asserts rank % 16 == 0: a rank-0 layer
    quantises without a low-rank projection.  Codes and scales must equal those of the rank-32 call on the same input."""
    from nunchaku_b200.ops.quantize import svdq_quantize_w4a4_act_fuse_lora_cuda
    from oracle import formats as F
    from oracle import svdq as O

    hT = torch.bfloat16
    M, K = 300, 512
    layer = O.make_synthetic_layer(128, K, 32, fp4=fp4, hT=hT, seed=401)
    x = O.make_activations(M, K, hT, seed=402, smooth=layer.smooth).cuda()
    smooth = F.pack_channel_vector(layer.smooth).cuda()
    q32, s32, _ = svdq_quantize_w4a4_act_fuse_lora_cuda(x, lora_down=F.pack_lowrank(layer.lora_down, down=True).cuda(), smooth=smooth, fp4=fp4)
    q0, s0, la0 = svdq_quantize_w4a4_act_fuse_lora_cuda(x, lora_down=torch.empty(K, 0, dtype=hT, device="cuda"), smooth=smooth, fp4=fp4)
    torch.cuda.synchronize()
    assert la0.shape == (512, 0)
    assert torch.equal(q0[:M], q32[:M]) and torch.equal(s0.view(torch.uint8), s32.view(torch.uint8))
