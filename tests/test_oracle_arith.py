"""Self-consistency of the CPU oracle (no GPU).  The arithmetic is parity-unpinned
(see oracle/svdq.py header); these tests pin its internal invariants:
  * correct single rounding helper,
  * e2m1 / e4m3 codecs against torch's own fp8 tables and hand-computed ties,
  * quantise -> dequantise error bounds (half a step per element),
  * ref-emulating GEMM stays within the reference's own noise of the fp64 GEMM,
  * the whole SVDQuant layer approximates the un-quantised linear it was built from.
"""
import math

import pytest
import torch

from oracle import svdq as O


def test_rn_single_rounding_bf16():
    # 1 + 2^-8 is an exact bf16 tie (-> 1.0, even); adding 2^-40 must round up.  A naive
    # double->float->bf16 conversion loses the 2^-40 and rounds down.
    x = torch.tensor([1.0 + 2.0 ** -8, 1.0 + 2.0 ** -8 + 2.0 ** -40, 1.0 + 3 * 2.0 ** -8], dtype=torch.float64)
    y = O.rn(x, torch.bfloat16).to(torch.float64)
    assert y.tolist() == [1.0, 1.0 + 2.0 ** -7, 1.0 + 2.0 ** -6]
    z = O.rn(torch.tensor([1.0 + 2.0 ** -11 + 2.0 ** -45], dtype=torch.float64), torch.float16).to(torch.float64)
    assert z.item() == 1.0 + 2.0 ** -10


def test_e2m1_codec():
    vals = torch.tensor([0.0, 0.25, 0.26, 0.75, 1.25, 1.75, 2.5, 2.51, 3.5, 5.0, 5.01, 100.0, float("nan"),
                         -0.25, -0.75, -7.0])
    codes = O.e2m1_encode(vals)
    dec = O.e2m1_decode(codes)
    assert dec.tolist() == [0.0, 0.0, 0.5, 1.0, 1.0, 2.0, 2.0, 3.0, 4.0, 4.0, 6.0, 6.0, 6.0, -0.0, -1.0, -6.0]
    allc = torch.arange(16, dtype=torch.int8)
    assert torch.equal(O.e2m1_encode(O.e2m1_decode(allc)) & 7, allc & 7)


def test_e4m3_codec():
    v = torch.tensor([0.0, 1.0, 1.0625, 448.0, 500.0, 2.0 ** -9, 0.3])
    b = O.e4m3_encode(v)
    d = O.e4m3_decode(b)
    assert d[0] == 0 and d[1] == 1 and d[3] == 448 and d[4] == 448 and d[5] == 2.0 ** -9
    assert d[2] in (1.0, 1.125)  # tie to even -> 1.0
    assert d[2] == 1.0


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
def test_quantize_roundtrip_bounds(fp4, hT):
    M, K, R = 100, 256, 16
    x = O.make_activations(M, K, hT, seed=3)
    smooth = (torch.rand(K) + 0.5).to(hT)
    ld = (torch.randn(R, K) * 0.05).to(hT)
    qa = O.quantize_w4a4_act_fuse_lora(x, smooth, ld, fp4=fp4)
    assert qa.q.shape == (256, K) and qa.lora_act.shape == (256, R)
    G = 16 if fp4 else 64
    assert qa.scales.shape == (K // G, 256)
    # padded rows are exactly zero after dequantisation
    deq = O.dequant(qa.q, qa.scales.t(), fp4)
    assert torch.all(deq[M:] == 0)
    assert torch.all(qa.lora_act[M:] == 0)
    xs = O.h_div(x, smooth.view(1, K)).to(torch.float64)
    err = (deq[:M] - xs).abs().view(M, K // G, G)
    s = (O.e4m3_decode(qa.scales) if fp4 else qa.scales.to(torch.float64)).t()[:M]
    # int4: half a step (+ hT rounding of the stored scale); fp4: half the widest e2m1 gap (2)
    step = (1.0 if fp4 else 0.5) * s.unsqueeze(-1) * (1.0 + 2.0 ** -7) + 1e-12
    # fp4 scale is rounded to e4m3 (up to 6.25 % off), values quantised with the unrounded one
    slack = 1.07 if fp4 else 1.0
    amax = xs.abs().view(M, K // G, G).amax(-1, keepdim=True)
    assert torch.all(err <= step * slack + amax * (0.07 if fp4 else 2.0 ** -7))
    # lora_act against plain fp64 matmul
    ref = x.to(torch.float64) @ ld.to(torch.float64).t()
    assert O.rel_fro(qa.lora_act[:M], ref) < 1e-6


def test_int4_quantizer_extremes():
    hT = torch.bfloat16
    x = torch.zeros(2, 64, dtype=hT)
    x[0, 0] = 7.0
    x[0, 1] = -7.0
    x[0, 2] = 3.5   # -> 3.5 ties to even 4
    x[0, 3] = 2.5   # -> 2
    qa = O.quantize_w4a4_act_fuse_lora(x, None, torch.zeros(16, 64, dtype=hT), fp4=False)
    assert qa.q[0, :4].tolist() == [7, -7, 4, 2]
    assert float(qa.scales[0, 0]) == 1.0
    assert torch.all(qa.q[1] == 0) and float(qa.scales[0, 1]) == 0.0   # all-zero group


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
def test_layer_ref_vs_exact(fp4, hT):
    N, K, R, M = 256, 512, 32, 130
    layer = O.make_synthetic_layer(N, K, R, fp4=fp4, hT=hT, seed=5)
    x = O.make_activations(M, K, hT, seed=6)
    y_ref = O.svdq_linear_forward(layer, x, mode="ref")
    y_ex = O.svdq_linear_forward(layer, x, mode="exact")
    assert y_ref.shape == (M, N) and y_ref.dtype == hT
    # reference-emulating chain vs exact arithmetic on identical quantised operands
    e = O.rel_fro(y_ref, y_ex)
    assert e < (2e-2 if hT == torch.bfloat16 else 3e-3), e


@pytest.mark.parametrize("fp4", [False, True])
def test_layer_approximates_dense_linear(fp4):
    """End-to-end sanity of the synthetic SVDQuant construction: W4A4 + low-rank ~= W x."""
    hT = torch.bfloat16
    N, K, R, M = 256, 512, 32, 64
    layer = O.make_synthetic_layer(N, K, R, fp4=fp4, hT=hT, seed=7, with_wcscales=False)
    x = O.make_activations(M, K, hT, seed=8, smooth=layer.smooth)
    y = O.svdq_linear_forward(layer, x, mode="exact").to(torch.float64)
    # rebuild the dense weight the layer encodes
    Wq = O.dequant(layer.qw, layer.wscales, fp4) * layer.alpha
    W = Wq / layer.smooth.to(torch.float64).view(1, K) + layer.lora_up.to(torch.float64) @ layer.lora_down.to(torch.float64)
    y_dense = x.to(torch.float64) @ W.t() + layer.bias.to(torch.float64)
    assert O.rel_fro(y, y_dense) < 0.15


def test_fused_gelu_quantize_next_layer():
    hT = torch.bfloat16
    N, K, R, M = 256, 256, 16, 40
    layer = O.make_synthetic_layer(N, K, R, fp4=False, hT=hT, seed=9)
    nxt_smooth = (torch.rand(N) + 0.5).to(hT)
    nxt_ld = (torch.randn(16, N) * 0.05).to(hT)
    x = O.make_activations(M, K, hT, seed=10)
    qa = O.quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down)
    res = O.gemm_w4a4(qa=qa.q, ascales=qa.scales, qw=layer.qw, wscales=layer.wscales, hT=hT, M=M,
                      bias=layer.bias, lora_act=qa.lora_act, lora_up=layer.lora_up, act="gelu",
                      next_smooth=nxt_smooth, next_lora_down=nxt_ld, want_out=True)
    assert res.qout.shape == (256, N) and res.oscales.shape == (N // 64, 256)
    assert int(res.qout.min()) >= 0 and int(res.qout.max()) <= 15
    # dequantised next-layer input ~ (gelu + shift)/smooth
    g = res.out.to(torch.float64)
    tgt = (g + O.SHIFT_GELU) / nxt_smooth.to(torch.float64).view(1, N)
    deq = O.dequant(res.qout, res.oscales.t(), False)[:M]
    assert O.rel_fro(deq, tgt) < 0.08
    assert O.rel_fro(res.lora_act_out[:M], g @ nxt_ld.to(torch.float64).t()) < 1e-5
    # gelu >= -0.17 so the shifted value is non-negative up to hT rounding
    assert float((g + O.SHIFT_GELU).min()) > -1e-2


def test_rmsnorm_rope_properties():
    torch.manual_seed(0)
    M, H = 8, 2
    N = 3 * H * 128
    y = torch.randn(M, N, dtype=torch.float64)
    wq = torch.ones(128, dtype=torch.bfloat16)
    wk = torch.ones(128, dtype=torch.bfloat16)
    ang = torch.rand(M, 64) * 6.28
    out = O.rmsnorm_rope(y, wq, wk, torch.sin(ang), torch.cos(ang), eps=0.0)
    # V untouched, Q/K heads have unit RMS (rotation preserves norm)
    assert torch.equal(out[:, 2 * N // 3:], y[:, 2 * N // 3:])
    rms = out[:, :2 * N // 3].reshape(M, 2 * H, 128).pow(2).mean(-1).sqrt()
    assert torch.allclose(rms, torch.ones_like(rms), atol=1e-9)
