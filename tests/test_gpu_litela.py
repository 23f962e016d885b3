"""Row a13: SANA linear-attention pieces (EpilogueLiteLA + vk_mul_q) vs oracle/glue.py, both routes: "fused" = the reduction inside the
GEMM epilogue (csrc/gemm_w4a4.cu EPI_LITELA: tensor-core Gram matrix of the staged K | V tile), "split" = plain GEMM + nb200_litela_vk.

Tolerances: relu(Q) bit-exact (it is the same hT GEMM output as the plain epilogue); vk fp32 atomics / summation
order vs fp64: <= 1e-5 relative to the row scale on the split route (sequential fp32 FMAs) and <= 3e-5 on the fused one (the tensor
core adds 16 exact products per instruction into an fp32 accumulator with truncation, not round-to-nearest; the reference's
mma.sync accumulation behaves the same way); vk_mul_q: fp32 FMA chain + div.approx vs fp64 + exact division,
then rounded to hT: <= 1 ulp of hT on <= 2 % of the elements."""
import pytest
import torch

import b200_layouts as L
from gpu_util import ref_layout_params
from oracle import glue as OG
from oracle import svdq as O

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fused", "split"])
def route(request, monkeypatch):
    import nunchaku_b200.ops.gemm as G

    monkeypatch.setattr(G, "LITELA_FUSED", request.param == "fused")
    return request.param


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
def test_litela_epilogue(fp4, hT, route):
    from nunchaku_b200._C import lib
    from nunchaku_b200.ops.gemm import linearattn_vk_mul_q, svdq_gemm_w4a4_cuda
    from test_gpu_fused import _pack_act

    heads, K, R, B, T = 4, 256, 32, 2, 256
    N = 3 * heads * 32
    layer = O.make_synthetic_layer(N, K, R, fp4=fp4, hT=hT, seed=101)
    x = O.make_activations(B * T, K, hT, seed=102, smooth=layer.smooth)
    qa = O.quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down, fp4=fp4)
    params = ref_layout_params(layer)
    act, asc = _pack_act(qa, fp4, hT)
    common = dict(act=act, wgt=params["qweight"], ascales=asc, wscales=params["wscales"], lora_act_in=qa.lora_act.cuda(),
                  lora_up=params["proj_up"], bias=params["bias"], fp4=fp4, alpha=layer.alpha, wcscales=params["wcscales"])
    plain = torch.empty(B * T, N, dtype=hT, device="cuda")
    svdq_gemm_w4a4_cuda(out=plain, **common)
    out_q = torch.full((B, T, N // 3), float("nan"), dtype=hT, device="cuda")
    out_vk = torch.full((B, heads, 33, 32), float("nan"), dtype=torch.float32, device="cuda")
    svdq_gemm_w4a4_cuda(out_vk=out_vk, out_linearattn=out_q, **common)
    torch.cuda.synchronize()
    if route == "fused":
        assert lib.nb200_last_launch_count() == 1   # ONE kernel (GEMM + reduction), no scratch
    want_q, want_vk = OG.litela_vk(plain.cpu().view(B, T, N))
    assert torch.equal(out_q.cpu().view(torch.int16), want_q.view(torch.int16))
    scale = want_vk.abs().amax(dim=-1, keepdim=True).clamp_min(1e-6)
    assert ((out_vk.cpu() - want_vk).abs() / scale).max().item() <= (3e-5 if route == "fused" else 1e-5)
    # second kernel, in place on q
    q2 = out_q.clone()
    linearattn_vk_mul_q(q2, out_vk)
    torch.cuda.synchronize()
    want = OG.vk_mul_q(out_q.cpu(), out_vk.cpu())
    bits = 7 if hT == torch.bfloat16 else 10
    a, b = q2.cpu().double(), want.double()
    ulp = torch.exp2(torch.floor(torch.log2(torch.maximum(a.abs(), b.abs()).clamp_min(2.0 ** -14))) - bits)
    d = (a - b).abs() / ulp
    assert d.max().item() <= 1.0 and (d > 0).double().mean().item() <= 0.02
    with pytest.raises(ValueError):
        svdq_gemm_w4a4_cuda(out_vk=out_vk, **common)


@pytest.mark.parametrize("fp4", [False, True])
def test_litela_fused_sana_shape(fp4):
    """SANA-1.6B's padded QKV projection (72 heads of 32 = 3 x 2304 channels, 2 images x 1024 tokens): 54 tiles per m-block, 16 m-blocks, more
    tiles than SMs (the persistent loop reuses the staging buffers and the Gram accumulator many times per CTA), both routes against each other
    and against the fp64 reduction of the plain projection."""
    import nunchaku_b200.ops.gemm as G
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda
    from test_gpu_fused import _pack_act

    hT = torch.bfloat16
    heads, K, R, B, T = 72, 256, 32, 2, 1024
    N = 3 * heads * 32
    layer = O.make_synthetic_layer(N, K, R, fp4=fp4, hT=hT, seed=111)
    x = O.make_activations(B * T, K, hT, seed=112, smooth=layer.smooth)
    qa = O.quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down, fp4=fp4)
    params = ref_layout_params(layer)
    act, asc = _pack_act(qa, fp4, hT)
    common = dict(act=act, wgt=params["qweight"], ascales=asc, wscales=params["wscales"], lora_act_in=qa.lora_act.cuda(),
                  lora_up=params["proj_up"], bias=params["bias"], fp4=fp4, alpha=layer.alpha, wcscales=params["wcscales"])
    plain = torch.empty(B * T, N, dtype=hT, device="cuda")
    svdq_gemm_w4a4_cuda(out=plain, **common)
    res = {}
    for fused in (True, False):
        G.LITELA_FUSED = fused
        try:
            out_q = torch.full((B, T, N // 3), float("nan"), dtype=hT, device="cuda")
            out_vk = torch.full((B, heads, 33, 32), float("nan"), dtype=torch.float32, device="cuda")
            for _ in range(2):   # twice: out_vk is zero-filled inside the call, not accumulated across calls
                svdq_gemm_w4a4_cuda(out_vk=out_vk, out_linearattn=out_q, **common)
            torch.cuda.synchronize()
            res[fused] = (out_q.cpu(), out_vk.cpu())
        finally:
            G.LITELA_FUSED = None
    want_q, want_vk = OG.litela_vk(plain.cpu().view(B, T, N))
    scale = want_vk.abs().amax(dim=-1, keepdim=True).clamp_min(1e-6)
    for fused, (got_q, got_vk) in res.items():
        assert torch.equal(got_q.view(torch.int16), want_q.view(torch.int16)), fused
        assert ((got_vk - want_vk).abs() / scale).max().item() <= (1e-4 if fused else 3e-5), fused   # 1024 tokens: 8 atomics per entry on top
