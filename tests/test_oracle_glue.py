"""CPU checks of oracle/glue.py against independent formulas (torch fp32/fp64)."""
import pytest
import torch

from oracle import glue as O


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_activations_close_to_fp64(dtype):
    g = torch.Generator().manual_seed(0)
    x = (3 * torch.randn(4096, generator=g)).to(dtype)
    eps = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    xs = x.double()
    silu = xs / (1 + torch.exp(-xs))
    assert ((O.silu(x).double() - silu).abs() <= eps * silu.abs() + 1e-7).all()
    gelu = torch.nn.functional.gelu(xs, approximate="tanh")
    # gelu_new rounds x^3, the tanh argument, tanh and (1 + tanh) to T: the error is absolute, ~eps * |x|
    # (in the negative tail 1 + T(tanh) cancels to 0 -- that IS the reference's behaviour)
    assert ((O.gelu_new(x).double() - gelu).abs() <= 2 * eps * xs.abs().clamp_min(1.0)).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_norms_close_to_torch(dtype):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(33, 3072, generator=g) * 2 + 0.3).to(dtype)
    w = (1 + 0.1 * torch.randn(3072, generator=g)).to(dtype)
    b = (0.1 * torch.randn(3072, generator=g)).to(dtype)
    eps = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    ref = torch.nn.functional.layer_norm(x.double(), (3072,), w.double(), b.double(), 1e-6)
    assert (O.layernorm(x, w, b, 1e-6).double() - ref).abs().max() <= eps * ref.abs().max()
    rms = x.double() * torch.rsqrt(x.double().square().mean(-1, keepdim=True) + 1e-6) * w.double()
    assert (O.rms_norm(x, w, 1e-6).double() - rms).abs().max() <= 2 * eps * rms.abs().max()


def test_mul_add_is_single_rounding_and_split_roundtrip():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 5, 64, generator=g).to(torch.bfloat16)
    s = torch.randn(2, 1, 64, generator=g).to(torch.bfloat16)
    b = torch.randn(1, 1, 64, generator=g).to(torch.bfloat16)
    out = O.mul_add_batch(x, s, True, 1.0, b, False)
    s1 = (s.double() + 1.0).to(torch.bfloat16).double()  # scale + shift rounds to T first
    exact = x.double() * s1 + b.double()
    assert ((out.double() - exact).abs() <= 2.0 ** -8 * exact.abs() + 1e-30).all()  # half an ulp: one rounding
    parts = O.split_mod(x, 4)
    assert torch.equal(torch.stack(parts, -1).reshape(x.shape), x)
    assert O.cast(torch.tensor([1e6, -1e6]), torch.float16).tolist() == [65504.0, -65504.0]


def test_litela_oracle_against_naive_loops():
    """oracle.glue.litela_vk / vk_mul_q vs the definition written as plain loops (epilogues.cuh:552-760)."""
    g = torch.Generator().manual_seed(3)
    B, T, heads = 2, 5, 2
    N = 3 * heads * 32
    qkv = torch.randn(B, T, N, generator=g).to(torch.bfloat16)
    q, vk = O.litela_vk(qkv)
    assert torch.equal(q, torch.clamp_min(qkv[..., : N // 3].float(), 0).to(torch.bfloat16))
    ref = torch.zeros(B, heads, 33, 32, dtype=torch.float64)
    for b in range(B):
        for h in range(heads):
            for t in range(T):
                k = torch.clamp_min(qkv[b, t, N // 3 + h * 64: N // 3 + h * 64 + 32].double(), 0)
                v = qkv[b, t, N // 3 + h * 64 + 32: N // 3 + h * 64 + 64].double()
                ref[b, h, :32] += v[:, None] * k[None, :]
                ref[b, h, 32] += k
    assert torch.allclose(vk.double(), ref, rtol=1e-6, atol=1e-6)
    out = O.vk_mul_q(q, vk)
    for b in range(B):
        for t in range(T):
            for h in range(heads):
                qv = q[b, t, h * 32:(h + 1) * 32].double()
                num = vk[b, h, :32].double() @ qv
                den = vk[b, h, 32].double() @ qv + 1e-6
                want = (num / den).float().to(torch.bfloat16)
                got = out[b, t, h * 32:(h + 1) * 32]
                assert (got.float() - want.float()).abs().max() <= 2.0 ** -7 * max(1.0, want.float().abs().max().item())
