"""GPU parity of the elementwise / row-reduce glue against oracle/glue.py.

Tolerances (written here as the task requires):
  * add, mul_add(_batch), split_mod, cast: bit-exact (the arithmetic is a fixed sequence of correctly
    rounded 16-bit ops; mul+add is one fused rounding on both sides, see oracle/glue.py).
  * silu, gelu_new: expf / tanhf are CUDA's fp32 library versions (<= 2 ulp fp32) vs glibc's on the CPU;
    after rounding to 16 bits at most a 1-ulp-of-T difference on a small fraction of elements.
  * layernorm, rms_norm: fp32 row statistics summed in a different order: mean / rstd can differ in the last
    fp32 bit, which moves an output by <= 1 ulp of T (layernorm; ulp taken at max(|y|, 2^-6) because
    (x - mean) cancels near zero) or <= 2 ulp of T (rms_norm rounds T(x * rstd) and then T(n * w)).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float16]


def _ulp_diff(a: torch.Tensor, b: torch.Tensor, floor: float = 0.0) -> torch.Tensor:
    """|a - b| in units of the larger operand's ulp (16-bit types); magnitudes below `floor` use floor's ulp."""
    bits = 7 if a.dtype == torch.bfloat16 else 10
    af, bf = a.double().cpu(), b.double().cpu()
    mag = torch.maximum(af.abs(), bf.abs()).clamp_min(max(floor, 2.0 ** -14 if a.dtype == torch.float16 else 2.0 ** -126))
    ulp = torch.exp2(torch.floor(torch.log2(mag)) - bits)
    return (af - bf).abs() / ulp


def _assert_ulp(got, want, max_ulp=1.0, max_frac=0.02, floor=0.0):
    d = _ulp_diff(got, want, floor)
    assert d.max().item() <= max_ulp, f"max diff {d.max().item()} ulp"
    assert (d > 0).double().mean().item() <= max_frac, f"{(d > 0).double().mean().item():.4f} of elements differ"


def _assert_equal(got, want):
    assert got.dtype == want.dtype and got.shape == want.shape
    assert torch.equal(got.cpu().view(torch.int16 if got.element_size() == 2 else torch.int32),
                       want.cpu().view(torch.int16 if want.element_size() == 2 else torch.int32))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(4352, 3072), (3, 1000), (1, 7)])
def test_activations(dtype, shape):
    from nunchaku_b200.ops import glue
    from oracle import glue as O

    g = torch.Generator().manual_seed(0)
    x = (3.0 * torch.randn(*shape, generator=g)).to(dtype)
    x.view(-1)[:6] = torch.tensor([0.0, -0.0, 20.0, -20.0, 1e-3, -1e-3], dtype=dtype)
    xd = x.cuda()
    _assert_ulp(glue.silu(xd), O.silu(x))
    _assert_ulp(glue.gelu_new(xd), O.gelu_new(x))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,hidden,affine", [(4352, 3072, False), (257, 3072, True), (5, 64, True), (3, 10240, True), (2, 20480, False)])
def test_layernorm(dtype, rows, hidden, affine):
    from nunchaku_b200.ops import glue
    from oracle import glue as O

    g = torch.Generator().manual_seed(1)
    x = (torch.randn(rows, hidden, generator=g) * 2 + 0.5).to(dtype)
    w = (1 + 0.1 * torch.randn(hidden, generator=g)).to(dtype) if affine else None
    b = (0.1 * torch.randn(hidden, generator=g)).to(dtype) if affine else None
    got = glue.layernorm(x.cuda(), None if w is None else w.cuda(), None if b is None else b.cuda(), 1e-6)
    _assert_ulp(got, O.layernorm(x, w, b, 1e-6), max_frac=0.05, floor=2.0 ** -6)
    # independent check against torch's own fp32 layer_norm
    ref = torch.nn.functional.layer_norm(x.float(), (hidden,), None if w is None else w.float(), None if b is None else b.float(), 1e-6)
    assert (got.float().cpu() - ref).abs().max().item() <= 2.0 ** (-7 if dtype == torch.bfloat16 else -10) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,hidden", [(4352 * 24, 128), (300, 3072), (2, 16384)])
def test_rms_norm(dtype, rows, hidden):
    from nunchaku_b200.ops import glue
    from oracle import glue as O

    g = torch.Generator().manual_seed(2)
    x = torch.randn(rows, hidden, generator=g).to(dtype)
    w = (1 + 0.1 * torch.randn(hidden, generator=g)).to(dtype)
    _assert_ulp(glue.rms_norm(x.cuda(), w.cuda(), 1e-6), O.rms_norm(x, w, 1e-6), max_ulp=2.0, max_frac=0.05)


@pytest.mark.parametrize("dtype", DTYPES + [torch.float32])
def test_add_bit_exact(dtype):
    from nunchaku_b200.ops import glue
    from oracle import glue as O

    g = torch.Generator().manual_seed(3)
    for n in (4352 * 3072, 1001, 8, 1):
        a = (torch.randn(n, generator=g) * 100).to(dtype)
        b = torch.randn(n, generator=g).to(dtype)
        _assert_equal(glue.add(a.cuda(), b.cuda()), O.add(a, b))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("with_scale,batch_scale,batch_bias,shift", [(True, True, True, 1.0), (True, False, True, 0.0), (False, False, True, 0.0),
                                                                      (True, True, False, 1.0)])
def test_mul_add_batch_bit_exact(dtype, with_scale, batch_scale, batch_bias, shift):
    """The AdaLN modulation call: x [B, T, C], scale / shift [B, 1, C] (FluxModel.cpp AdaLayerNorm)."""
    from nunchaku_b200.ops import glue
    from oracle import glue as O

    g = torch.Generator().manual_seed(4)
    B, T, C = 3, 257, 3072
    x = torch.randn(B, T, C, generator=g).to(dtype)
    scale = torch.randn(B if batch_scale else 1, 1, C, generator=g).to(dtype) if with_scale else None
    bias = torch.randn(B if batch_bias else 1, 1, C, generator=g).to(dtype)
    if dtype == torch.float16:
        x[0, 0, :8] = 60000.0  # exercises the +-65504 clamp
    want = O.mul_add_batch(x, scale, batch_scale, shift, bias, batch_bias)
    xd = x.cuda()
    glue.mul_add_batch(xd, None if scale is None else scale.cuda(), batch_scale, shift, bias.cuda(), batch_bias)
    _assert_equal(xd, want)
    assert torch.isfinite(xd.float()).all()


@pytest.mark.parametrize("dtype", DTYPES)
def test_mul_add_bit_exact(dtype):
    from nunchaku_b200.ops import glue
    from oracle import glue as O

    g = torch.Generator().manual_seed(5)
    x = torch.randn(512, 3072, generator=g).to(dtype)
    scale = torch.randn(3072, generator=g).to(dtype)
    bias = torch.randn(3072, generator=g).to(dtype)
    want = O.mul_add_batch(x.unsqueeze(0), scale, False, 0.0, bias, False)[0]
    xd = x.cuda()
    glue.mul_add(xd, scale.cuda(), bias.cuda())
    _assert_equal(xd, want)


@pytest.mark.parametrize("dtype", DTYPES + [torch.float32])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6])
def test_split_mod_bit_exact(dtype, n):
    from nunchaku_b200.ops import glue
    from oracle import glue as O

    g = torch.Generator().manual_seed(6)
    for lead, c in (((2, 1), 3072), ((7,), 10)):
        x = torch.randn(*lead, c * n, generator=g).to(dtype)
        got = glue.split_mod(x.cuda(), n)
        for a, b in zip(got, O.split_mod(x, n)):
            _assert_equal(a, b)


@pytest.mark.parametrize("src", DTYPES + [torch.float32])
@pytest.mark.parametrize("dst", DTYPES + [torch.float32])
def test_cast_bit_exact(src, dst):
    from nunchaku_b200.ops import glue
    from oracle import glue as O

    g = torch.Generator().manual_seed(7)
    x = (torch.randn(1003, 33, generator=g) * 1000).to(src)
    x.view(-1)[:2] = torch.tensor([1e30 if src != torch.float16 else 65504.0, -1e30 if src != torch.float16 else -65504.0]).to(src)
    out = torch.empty(x.shape, dtype=dst, device="cuda")
    glue.cast(x.cuda(), out)
    _assert_equal(out, O.cast(x, dst))


def test_preconditions():
    from nunchaku_b200.ops import glue

    x = torch.randn(4, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        glue.silu(x.cpu())
    with pytest.raises(TypeError):
        glue.silu(x.to(torch.float64))
    with pytest.raises(ValueError):
        glue.split_mod(x, 7)
    with pytest.raises(ValueError):
        glue.add(x, x[:2])
    with pytest.raises(RuntimeError):
        glue.layernorm(torch.randn(4, 12, device="cuda", dtype=torch.bfloat16))  # hidden % 8 != 0


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,hidden", [(300, 3072), (17, 2240), (5, 8192)])
def test_layernorm_mod_equals_layernorm_then_mul_add(dt, rows, hidden):
    """AdaLN in one pass (nb200_layernorm_mod) == layernorm followed by mul_add_batch, bit for bit"""
    from nunchaku_b200.ops import glue

    g = torch.Generator(device="cuda").manual_seed(rows)
    x = (torch.randn(1, rows, hidden, generator=g, device="cuda") * 3 + 0.5).to(dt)
    scale = (0.3 * torch.randn(1, hidden, generator=g, device="cuda")).to(dt)
    shift = (0.3 * torch.randn(1, hidden, generator=g, device="cuda")).to(dt)
    want = glue.layernorm(x, None, None, 1e-6)
    glue.mul_add_batch(want, scale, True, 1.0, shift, True)
    got = glue.layernorm_mod(x, scale, shift, 1e-6)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
