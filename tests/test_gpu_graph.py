"""GraphedStep (nunchaku_b200/graph.py): a captured stack of SVDQuant layers replays bit-identically to eager launches,
also after the inputs changed."""
import pytest
import torch

from oracle import svdq as O

pytestmark = pytest.mark.gpu


def _layer(K, N, precision, seed, unsigned=False):
    from gpu_util import ref_layout_params
    from nunchaku_b200.models.linear import SVDQW4A4Linear

    fp4 = precision == "nvfp4"
    layer = O.make_synthetic_layer(N, K, 32, fp4=fp4, hT=torch.bfloat16, seed=seed)
    p = ref_layout_params(layer)
    m = SVDQW4A4Linear(K, N, rank=32, bias=True, precision=precision, act_unsigned=unsigned, torch_dtype=torch.bfloat16, device="cuda")
    sd = {"qweight": p["qweight"], "wscales": p["wscales"], "bias": p["bias"], "smooth_factor": p["smooth"], "smooth_factor_orig": p["smooth"],
          "proj_down": p["proj_down"], "proj_up": p["proj_up"]}
    if fp4:
        sd["wcscales"] = p["wcscales"]
        m.wtscale = layer.alpha
    m.load_state_dict(sd)
    return m


@pytest.mark.parametrize("precision", ["nvfp4", "int4"])
def test_graphed_block_matches_eager(precision):
    from nunchaku_b200.graph import GraphedStep
    from nunchaku_b200.ops import glue

    D, H, M = 256, 512, 300
    fc1 = _layer(D, H, precision, 301)
    fc2 = _layer(H, D, precision, 302, unsigned=precision == "int4")
    proj = _layer(D, D, precision, 303)

    def block(x):
        n = glue.layernorm(x, None, None, 1e-6)
        a = proj(n.view(1, M, D)).view(M, D)
        f = fc1.forward_mlp(n, fc2, fuse=True)    # (bit-stable: the fused hand-off reduces its low-rank partials in a fixed order)
        return glue.add(glue.add(x, a), f)

    g = torch.Generator(device="cuda").manual_seed(5)
    x0 = torch.randn(M, D, generator=g, device="cuda").to(torch.bfloat16)
    x1 = torch.randn(M, D, generator=g, device="cuda").to(torch.bfloat16)
    step = GraphedStep(block, (x0,))
    for x in (x0, x1, x0):
        want = block(x).clone()
        got = step(x)
        torch.cuda.synchronize()
        assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    with pytest.raises(ValueError):
        step(x0[:100])
