"""GPU parity against the REFERENCE'S OWN KERNELS running on the same B200 (INT4; the reference's NVFP4 kernels
cannot execute on sm_100), at BASELINE.json's full FLUX shapes, three ways:

  ours   nunchaku_b200 Python operator layer -> C ABI -> tcgen05 kernels
  ref    oracle/_ref/libnunchaku_ref.so: the reference's unmodified src/Linear.cpp + src/kernels/zgemm/* built for sm_100a
  seam   oracle/_ref/libnunchaku_seam.so: the reference's unmodified src/Linear.cpp (class GEMM_W4A4), Module.cpp,
         activation.cpp, layernorm.cpp linked on top of OUR definitions of zgemm.h / misc_kernels.h
         (nunchaku_b200/csrc/seam/*.cpp) -- SURVEY section 8 rows a4 and (b): the C++ drop-in

on identical checkpoint tensors and inputs.  "exact" is an fp64 evaluation on the GPU of the oracle-quantised
operands.  The reference accumulates in 16-bit (bf16: ~7e-3 of noise at K=3072, more at K=12288), ours in fp32, so
  ours-vs-exact <= ref-vs-exact            (we are at least as accurate as the reference)
  ours-vs-ref   <= max(1e-2, 1.25 * ref-vs-exact)   (north_star's 1e-2, widened only where the reference's own
                                                     distance from exact math already exceeds it)
Both libraries are built by oracle/ref_build/build_ref.sh in the authoring container and travel to the GPU box.
"""
import json
import os

import pytest
import torch

from gpu_util import diag
from oracle import formats as F
from oracle import refgpu as R
from oracle import svdq as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available("ref"), reason="oracle/_ref/libnunchaku_ref.so not built")]
DEV = "cuda"


def _record(name, **kw):
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_vs_reference.jsonl"), "a") as f:
            f.write(json.dumps({"name": name, **kw}) + "\n")
    except OSError:
        pass


def fast_layer(N, K, Rk, hT, seed):
    """a quantised layer with random codes (no SVD: generated on the GPU in milliseconds), logical form"""
    g = torch.Generator(device=DEV).manual_seed(seed)
    qw = torch.randint(-8, 8, (N, K), generator=g, device=DEV, dtype=torch.int8)
    wscales = (torch.rand(N, K // 64, generator=g, device=DEV) * 0.004 + 0.001).to(hT)
    bias = (torch.randn(N, generator=g, device=DEV) * 0.1).to(hT)
    smooth = torch.exp(torch.randn(K, generator=g, device=DEV) * 0.5).clamp(0.1, 10).to(hT)
    lora_up = (torch.randn(N, Rk, generator=g, device=DEV) * 0.05).to(hT)
    lora_down = (torch.randn(Rk, K, generator=g, device=DEV) * 0.05 / smooth.float().view(1, K)).to(hT)
    return O.SynthLayer(qw=qw, wscales=wscales, bias=bias, smooth=smooth, lora_down=lora_down, lora_up=lora_up, wcscales=None, alpha=1.0,
                        fp4=False, hT=hT)


def checkpoint(layer):
    return {
        "qweight": F.pack_qweight(layer.qw), "wscales": F.pack_group_scales(layer.wscales), "bias": F.pack_channel_vector(layer.bias),
        "smooth": F.pack_channel_vector(layer.smooth), "lora_up": F.pack_lowrank(layer.lora_up, down=False),
        "lora_down": F.pack_lowrank(layer.lora_down, down=True),
    }


def ref_module(layer, ck, lib):
    N, K = layer.qw.shape
    return R.RefLinear(K, N, bias=True, fp4=False, dtype=layer.hT, lib=lib).load(**ck)


def our_module(layer, ck, act_unsigned=False):
    from nunchaku_b200.models.linear import SVDQW4A4Linear

    N, K = layer.qw.shape
    m = SVDQW4A4Linear(K, N, rank=layer.lora_up.shape[1], bias=True, precision="int4", act_unsigned=act_unsigned, torch_dtype=layer.hT, device=DEV)
    m.load_state_dict({"qweight": ck["qweight"], "wscales": ck["wscales"], "bias": ck["bias"], "smooth_factor": ck["smooth"],
                       "smooth_factor_orig": ck["smooth"], "proj_down": ck["lora_down"], "proj_up": ck["lora_up"]})
    return m


def exact_linear(layer, x):
    """fp64 on the GPU from oracle-quantised activations (exact division; the kernels use rcp.approx: a few 1e-4 of the
    codes sit one step away, far below the gates)"""
    K = x.shape[1]
    xs = (x.float() / layer.smooth.float().view(1, K)).to(layer.hT)
    g = xs.view(x.shape[0], K // 64, 64)
    s32 = g.abs().amax(-1).float() * (1.0 / 7.0)
    sc = s32.to(layer.hT)
    q = torch.round(g.float() * (1.0 / s32.double()).float().unsqueeze(-1)).nan_to_num(0.0).clamp(-8, 7)
    a = (q.double() * sc.double().unsqueeze(-1)).view(x.shape[0], K)
    w = (layer.qw.double().view(-1, K // 64, 64) * layer.wscales.double().unsqueeze(-1)).view(-1, K)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y = a @ w.t() + layer.bias.double().view(1, -1)
        la = x.double() @ layer.lora_down.double().t()
        y = y + la.float().to(layer.hT).double() @ layer.lora_up.double().t()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return y


def gates(name, ours, ref, seam, exact, hT):
    e = lambda a, b: O.rel_fro(a.cpu(), b.cpu())  # noqa: E731
    n = {"ours_vs_ref": e(ours, ref), "seam_vs_ref": e(seam, ref), "seam_vs_ours": e(seam, ours)}
    if exact is not None:
        n.update(ours_vs_exact=e(ours, exact), ref_vs_exact=e(ref, exact))
    _record(name, dtype=str(hT), **n)
    msg = f"{name}: {n}\n" + diag(name, ours, ref)
    assert not torch.isnan(ours).any() and not torch.isnan(seam).any(), msg
    assert n["seam_vs_ours"] <= 1e-6, msg                       # same kernels behind a different host layer
    if exact is not None:
        assert n["ours_vs_exact"] <= n["ref_vs_exact"] + 1e-3, msg
        assert n["ours_vs_ref"] <= max(1e-2, 1.25 * n["ref_vs_exact"]), msg
    return n


SHAPES = [  # (M, K, N) of the FLUX.1 linears: out-proj, qkv-sized, fc2 (CTA-pair auto-dispatch), text-stream fc1
    (4352, 3072, 3072), (4352, 3072, 9216), (4352, 12288, 3072), (256, 3072, 12288), (300, 3072, 3072),
]


@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,K,N", SHAPES)
def test_linear_forward_vs_reference_kernels(M, K, N, hT):
    if hT == torch.float16 and (M, K, N) not in [(4352, 3072, 3072), (4352, 12288, 3072)]:
        pytest.skip("fp16: two shapes are enough")
    layer = fast_layer(N, K, 32, hT, seed=K + N)
    ck = checkpoint(layer)
    x = O.make_activations(M, K, hT, seed=5, smooth=layer.smooth.cpu()).to(DEV)
    ours = our_module(layer, ck)(x.view(1, M, K)).view(M, N)
    ref = ref_module(layer, ck, "ref").forward(x)
    seam = ref_module(layer, ck, "seam").forward(x)
    torch.cuda.synchronize()
    gates(f"linear {M}x{K}->{N}", ours, ref, seam, exact_linear(layer, x), hT)


@pytest.mark.parametrize("M", [256, 4352])
def test_fused_gelu_mlp_vs_reference_kernels(M):
    """fc1 -> GELU -> quantise(next) + low-rank down(next) -> fc2 (unsigned activations): FluxModel.cpp's MLP."""
    from nunchaku_b200.ops.fused import fused_gelu_mlp

    hT = torch.bfloat16
    K, H = 3072, 12288
    l1, l2 = fast_layer(H, K, 32, hT, seed=1), fast_layer(K, H, 32, hT, seed=2)
    c1, c2 = checkpoint(l1), checkpoint(l2)
    x = O.make_activations(M, K, hT, seed=6, smooth=l1.smooth.cpu()).to(DEV)
    ours = fused_gelu_mlp(x.view(1, M, K), our_module(l1, c1), our_module(l2, c2, act_unsigned=True)).view(M, K)
    ref = ref_module(l1, c1, "ref").forward_mlp(ref_module(l2, c2, "ref"), x)
    seam = ref_module(l1, c1, "seam").forward_mlp(ref_module(l2, c2, "seam"), x)
    torch.cuda.synchronize()
    e = lambda a, b: O.rel_fro(a.cpu(), b.cpu())  # noqa: E731
    n = {"ours_vs_ref": e(ours, ref), "seam_vs_ref": e(seam, ref), "seam_vs_ours": e(seam, ours)}
    _record(f"fused mlp M={M}", **n)
    # two quantisers in the chain (hidden activations re-quantised to 4 bit): the reference's 16-bit noise on fc1's output
    # moves a fraction of the hidden codes by one step, and K = 12288 in fc2 accumulates 192 bf16 roundings
    assert n["ours_vs_ref"] <= 2.5e-2 and n["seam_vs_ref"] <= 2.5e-2, f"{n}\n" + diag("mlp", ours, ref)


@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
def test_qkv_rmsnorm_rope_vs_reference_kernels(hT):
    """QKV projection, 24 heads, RMSNorm(Q, K) + RoPE in the epilogue with a real pack_rotemb table (FluxModel.cpp:504-526)."""
    from nunchaku_b200.ops.fused import fused_qkv_norm_rottary

    M, K, N = 4352, 3072, 9216
    layer = fast_layer(N, K, 32, hT, seed=3)
    ck = checkpoint(layer)
    x = O.make_activations(M, K, hT, seed=7, smooth=layer.smooth.cpu()).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(8)
    norm_q = (1 + 0.1 * torch.randn(128, generator=g, device=DEV)).to(hT)
    norm_k = (1 + 0.1 * torch.randn(128, generator=g, device=DEV)).to(hT)
    theta = torch.rand(M, 64, generator=g, device=DEV) * 6.28318
    rot = F.pack_rotemb(torch.sin(theta).float(), torch.cos(theta).float()).view(1, M, 128).contiguous()

    class W:  # the reference passes torch.nn.RMSNorm modules
        def __init__(self, w):
            self.weight = w

    ours = fused_qkv_norm_rottary(x.view(1, M, K), our_module(layer, ck), W(norm_q), W(norm_k), rot.view(M, 128)).view(M, N)
    ref = ref_module(layer, ck, "ref").forward_qkv(x, norm_q, norm_k, rot)
    seam = ref_module(layer, ck, "seam").forward_qkv(x, norm_q, norm_k, rot)
    torch.cuda.synchronize()
    e = lambda a, b: O.rel_fro(a.cpu(), b.cpu())  # noqa: E731
    n = {"ours_vs_ref": e(ours, ref), "seam_vs_ref": e(seam, ref), "seam_vs_ours": e(seam, ours),
         "v_part_ours_vs_ref": e(ours[:, 6144:], ref[:, 6144:])}
    _record(f"qkv rope {hT}", **n)
    assert n["seam_vs_ours"] <= 1e-6 and n["ours_vs_ref"] <= 1e-2, f"{n}\n" + diag("qkv", ours, ref)


def test_seam_glue_vs_reference_kernels():
    """the reference's Silu/GELU/LayerNorm/RMSNorm host classes and misc kernels entry points, on our kernels vs theirs"""
    g = torch.Generator(device=DEV).manual_seed(9)
    for hT in (torch.bfloat16, torch.float16):
        x = (torch.randn(2, 512, 3072, generator=g, device=DEV) * 2).to(hT)
        y = torch.randn(2, 512, 3072, generator=g, device=DEV).to(hT)
        w = (1 + 0.1 * torch.randn(3072, generator=g, device=DEV)).to(hT)
        b = (0.1 * torch.randn(3072, generator=g, device=DEV)).to(hT)
        sc = torch.randn(2, 1, 3072, generator=g, device=DEV).to(hT)
        pairs = {
            "silu": lambda L: R.glue_activation("silu", x, lib=L), "gelu": lambda L: R.glue_activation("gelu", x, lib=L),
            "layernorm": lambda L: R.glue_layernorm(x, w, b, 1e-6, lib=L), "layernorm_plain": lambda L: R.glue_layernorm(x, None, None, 1e-6, lib=L),
            "rms_norm": lambda L: R.glue_rms_norm(x, w, 1e-6, lib=L), "add": lambda L: R.glue_add(x, y, lib=L),
            "mul_add_batch": lambda L: R.glue_mul_add_batch(x.clone(), sc, True, 1.0, y[:, :1].contiguous(), True, lib=L),
            "cast": lambda L: R.glue_cast(x, torch.float32, lib=L),
        }
        for name, fn in pairs.items():
            a, r = fn("seam"), fn("ref")
            torch.cuda.synchronize()
            eq = float((a == r).double().mean())
            assert eq >= 0.999 and O.rel_fro(a.cpu(), r.cpu()) <= 2e-5, (name, hT, eq)
        outs_a, outs_r = R.glue_split_mod(x[:, :1, : 6 * 256].contiguous(), 6, lib="seam"), R.glue_split_mod(x[:, :1, : 6 * 256].contiguous(), 6, lib="ref")
        torch.cuda.synchronize()
        assert all(torch.equal(a, r) for a, r in zip(outs_a, outs_r))


def test_seam_reload_invalidates_converted_weights():
    """load -> forward -> load different weights into the SAME module -> forward: the converted copies must follow"""
    hT = torch.bfloat16
    M, K, N = 256, 256, 256
    la, lb = fast_layer(N, K, 32, hT, seed=11), fast_layer(N, K, 32, hT, seed=12)
    x = O.make_activations(M, K, hT, seed=13).to(DEV)
    m = ref_module(la, checkpoint(la), "seam")
    y_a = m.forward(x)
    m.load(**checkpoint(lb))
    y_b = m.forward(x)
    want_b = ref_module(lb, checkpoint(lb), "ref").forward(x)
    torch.cuda.synchronize()
    assert O.rel_fro(y_b.cpu(), want_b.cpu()) <= 1e-2 and O.rel_fro(y_a.cpu(), want_b.cpu()) > 0.5
