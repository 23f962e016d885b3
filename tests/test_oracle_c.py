"""CPU: the plain-C oracle (oracle/svdq_ref.c) against the Python restatement (oracle/svdq.py, mode="ref").

Both follow the same CUDA source line by line; the only freedom is the order of a few fp64 dot products (BLAS vs a plain
loop), which can move a result by at most one hT ulp at an exact rounding tie.  Gate: bit-identical on >= 99.9 % of the
outputs, never more than 1 ulp apart."""
import pytest
import torch

from oracle import csvdq
from oracle import svdq as O


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
def test_c_oracle_matches_python_oracle(fp4, hT):
    N, K, R, M = 256, 384 if not fp4 else 320, 32, 70
    K = 384 if not fp4 else 320
    layer = O.make_synthetic_layer(N, K, R, fp4=fp4, hT=hT, seed=201)
    x = O.make_activations(M, K, hT, seed=202, smooth=layer.smooth)
    want = O.svdq_linear_forward(layer, x, mode="ref")
    got = csvdq.linear_forward(layer, x)
    assert got.shape == want.shape and got.dtype == want.dtype
    bits = 7 if hT == torch.bfloat16 else 10
    a, b = got.double(), want.double()
    ulp = torch.exp2(torch.floor(torch.log2(torch.maximum(a.abs(), b.abs()).clamp_min(2.0 ** -14))) - bits)
    d = (a - b).abs() / ulp
    assert d.max().item() <= 1.0, d.max().item()
    assert (d > 0).double().mean().item() <= 1e-3, (d > 0).double().mean().item()


def test_c_oracle_zero_rows_and_ragged_m():
    layer = O.make_synthetic_layer(128, 128, 16, fp4=False, hT=torch.bfloat16, seed=203)
    x = O.make_activations(3, 128, torch.bfloat16, seed=204, smooth=layer.smooth)
    x[1] = 0                                                            # an all-zero row: scale 0, rcp = inf, codes 0
    want = O.svdq_linear_forward(layer, x, mode="ref")
    got = csvdq.linear_forward(layer, x)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
