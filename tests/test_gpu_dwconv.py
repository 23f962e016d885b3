"""Depthwise 3x3 convolution (csrc/dwconv.cu) against exact math (oracle/dwconv.py), SANA's shape and the edges."""
import pytest
import torch

from oracle import dwconv as DW
from oracle import svdq as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 32, 32, 11200), (1, 1, 1, 8), (1, 5, 1, 24), (3, 1, 7, 64), (1, 9, 13, 136)])
@pytest.mark.parametrize("use_bias", [True, False])
def test_dwconv_matches_exact_math(dt, shape, use_bias):
    from nunchaku_b200.ops.dwconv import DWCONV, dwconv_f16

    N, H, W, C = shape
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(N, H, W, C, generator=g).to(dt)
    w = (torch.randn(C, 3, 3, 1, generator=g) / 3).to(dt)
    b = torch.randn(C, generator=g).to(dt) if use_bias else None
    want = DW.dwconv3x3(x, w, b)
    got = dwconv_f16(x.cuda(), w.cuda(), None, None if b is None else b.cuda())
    torch.cuda.synchronize()
    # one rounding of an fp32 sum: half an ulp of hT per element
    assert O.rel_fro(got.cpu().double(), want) <= (3e-3 if dt == torch.bfloat16 else 4e-4)
    assert (got.cpu().double() - want).abs().max() <= (2 ** -7 if dt == torch.bfloat16 else 2 ** -10) * want.abs().max().clamp(min=1.0)
    m = DWCONV(C, use_bias, dt, "cuda")
    with torch.no_grad():
        m.weight.copy_(w)
        if use_bias:
            m.bias.copy_(b)
    assert torch.equal(m(x.cuda()), got)


def test_dwconv_argument_errors():
    from nunchaku_b200.ops.dwconv import dwconv_f16

    x = torch.zeros(1, 4, 4, 16, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(16, 3, 3, 1, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        dwconv_f16(x, w[:8].contiguous())
    with pytest.raises(ValueError):
        dwconv_f16(x, w, out=x)
    with pytest.raises(ValueError):
        dwconv_f16(x[..., :12].contiguous(), w[:12].contiguous())
    with pytest.raises(RuntimeError):
        dwconv_f16(x.cpu(), w.cpu())
