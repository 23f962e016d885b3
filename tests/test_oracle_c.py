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


@pytest.mark.parametrize("fp4", [False, True])
def test_baseline_config0_single_linear_3072(fp4):
    """BASELINE.json configs[0]: one SVDQuant linear 3072 x 3072, rank 32, on the CPU.  The reference-emulating C path, the
    reference-emulating Python path and the fp64 "exact" path on the same synthetic layer: C == Python (<= 1 ulp), and both sit
    within the reference's own arithmetic noise of the exact result (measured 7.5e-3 for the bf16 INT4 accumulation chain,
    well under 1e-2 = north_star's tolerance)."""
    hT = torch.bfloat16
    layer = O.make_synthetic_layer(3072, 3072, 32, fp4=fp4, hT=hT, seed=0)
    x = O.make_activations(64, 3072, hT, seed=1, smooth=layer.smooth)      # 64 rows keep the Python paths to a few seconds
    c = csvdq.linear_forward(layer, x)
    ref = O.svdq_linear_forward(layer, x, mode="ref")
    exact = O.svdq_linear_forward(layer, x, mode="exact")
    a, b = c.double(), ref.double()
    ulp = torch.exp2(torch.floor(torch.log2(torch.maximum(a.abs(), b.abs()).clamp_min(2.0 ** -14))) - 7)
    assert ((a - b).abs() / ulp).max().item() <= 1.0
    assert O.rel_fro(c, ref) <= 1e-4
    assert O.rel_fro(ref, exact) <= 1e-2 and O.rel_fro(c, exact) <= 1e-2
