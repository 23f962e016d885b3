"""GPU: nb200_repack_* kernels against the closed-form formats (bit-exact)."""
import pytest
import torch

import b200_layouts as L
from oracle import formats as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rp():
    from nunchaku_b200 import repack
    from nunchaku_b200._C import check, lib

    check(lib.nb200_check_device(), "check_device")
    return repack


def test_qweight_int4_and_fp4(rp):
    g = torch.Generator().manual_seed(0)
    N, K = 256, 384
    w = torch.randint(-8, 8, (N, K), generator=g, dtype=torch.int8)
    packed = F.pack_qweight(w).cuda()
    got = rp.qweight(packed, fp4=False).cpu()
    assert torch.equal(got, L.pack_int4(w, signed=True))
    codes = (w.to(torch.int16) & 0xF).to(torch.int8)
    packed4 = F.pack_qweight(codes).cuda()
    got4 = rp.qweight(packed4, fp4=True).cpu()
    assert torch.equal(got4, L.pack_fp4(codes))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_scales_and_vectors(rp, dtype):
    g = torch.Generator().manual_seed(1)
    N, K = 384, 512
    s = (torch.rand(N, K // 64, generator=g) + 0.5).to(dtype)
    got = rp.wscales(F.pack_group_scales(s).cuda(), N, K, fp4=False).cpu()
    assert torch.equal(got, s.t().contiguous())
    ms = torch.randint(1, 120, (N, K // 16), generator=g, dtype=torch.uint8)
    gotm = rp.wscales(F.pack_micro_scales(ms).view(torch.float8_e4m3fn).cuda(), N, K, fp4=True).cpu()
    assert torch.equal(gotm, L.pack_sf_tiles(ms))
    v = torch.randn(N, generator=g).to(dtype)
    pv = F.pack_channel_vector(v).cuda()
    assert torch.equal(rp.channel_vector(pv, out_f32=False).cpu(), v)
    assert torch.equal(rp.channel_vector(pv, out_f32=True, mul=0.5).cpu(), v.float() * 0.5)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("R", [16, 32, 48])
def test_lowrank(rp, dtype, R):
    g = torch.Generator().manual_seed(2)
    N, K = 256, 384
    lu = torch.randn(N, R, generator=g).to(dtype)
    ld = torch.randn(R, K, generator=g).to(dtype)
    cs = (torch.rand(N, generator=g) + 0.5)
    got_up = rp.lora_up(F.pack_lowrank(lu, down=False).cuda(), None).cpu()
    assert torch.equal(got_up, L.lora_up_blocks(lu))
    got_up2 = rp.lora_up(F.pack_lowrank(lu, down=False).cuda(), cs.cuda()).cpu()
    exp2 = L.lora_up_blocks(lu, cs)
    # fp32 division on device vs torch: allow 1 ulp of hT
    assert torch.allclose(got_up2.float(), exp2.float(), rtol=2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10, atol=0)
    got_dn = rp.lora_down(F.pack_lowrank(ld, down=True).cuda()).cpu()
    assert torch.equal(got_dn, L.lora_down_frags(ld))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_lora_up_division_by_the_channel_scale_is_validated_at_load_time(rp, dtype):
    """the epilogue keeps ONE accumulator, so lora_up is stored divided by alpha * wcscales; a scale the quotient cannot survive
    (zero / denormal / non-finite channel, fp16 overflow) is refused at repack time instead of producing inf / NaN outputs"""
    g = torch.Generator().manual_seed(3)
    N, R = 256, 32
    lu = (0.05 * torch.randn(N, R, generator=g)).to(dtype)
    packed = F.pack_lowrank(lu, down=False).cuda()
    small = torch.full((N,), 2.0e-4)                       # a realistic NVFP4 weight-tensor scale: quotients ~ 1e2 .. 1e3, fine in fp16
    got = rp.lora_up(packed, small.cuda(), cache=False).cpu()
    assert torch.isfinite(got.float()).all()
    assert torch.allclose(got.float(), L.lora_up_blocks(lu, small).float(), rtol=2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10, atol=0)
    zero = small.clone()
    zero[17] = 0.0
    with pytest.raises(RuntimeError, match="channel 17"):
        rp.lora_up(packed, zero.cuda(), cache=False)
    nan = small.clone()
    nan[200] = float("nan")
    with pytest.raises(RuntimeError, match="channel 200"):
        rp.lora_up(packed, nan.cuda(), cache=False)
    tiny = torch.full((N,), 1.0e-8)                        # 0.05 / 1e-8 = 5e6: beyond fp16 (65504), representable in bf16
    if dtype == torch.float16:
        with pytest.raises(RuntimeError, match="overflows"):
            rp.lora_up(packed, tiny.cuda(), cache=False)
    else:
        assert torch.isfinite(rp.lora_up(packed, tiny.cuda(), cache=False).float()).all()


@pytest.mark.parametrize("hT", [torch.float16, torch.bfloat16])
def test_nvfp4_layer_with_small_weight_tensor_scale_matches_the_oracle(hT):
    """fp16 + alpha ~ 1e-3 (what real NVFP4 checkpoints carry): lora_up / (alpha * wcscales) stays in range and the output matches"""
    from oracle import svdq as O

    layer = O.make_synthetic_layer(256, 384, 32, fp4=True, hT=hT, seed=311)
    assert layer.alpha < 5e-3
    m, _ = _module(layer, precision="nvfp4")
    x = O.make_activations(300, 384, hT, seed=312, smooth=layer.smooth)
    y = m(x.cuda().view(1, 300, 384)).view(300, 256)
    torch.cuda.synchronize()
    assert O.rel_fro(y.cpu(), O.svdq_linear_forward(layer, x, mode="ref")) <= (1e-2 if hT == torch.bfloat16 else 3e-3)


def _module(layer, precision="int4", act_unsigned=False, device="cuda"):
    from gpu_util import ref_layout_params
    from nunchaku_b200.models.linear import SVDQW4A4Linear

    N, K = layer.qw.shape
    p = ref_layout_params(layer, device=device)
    m = SVDQW4A4Linear(K, N, rank=layer.lora_up.shape[1], bias=True, precision=precision, act_unsigned=act_unsigned, torch_dtype=layer.hT, device=device)
    sd = {"qweight": p["qweight"], "wscales": p["wscales"], "bias": p["bias"], "smooth_factor": p["smooth"], "smooth_factor_orig": p["smooth"],
          "proj_down": p["proj_down"], "proj_up": p["proj_up"]}
    if layer.fp4:
        sd["wcscales"] = p["wcscales"]
        m.wtscale = layer.alpha
    m.load_state_dict(sd)
    return m, sd


@pytest.mark.parametrize("fp4", [False, True])
def test_module_owned_weights_follow_reloads_and_release(fp4):
    """load A -> forward -> load B into the SAME module -> forward must use B (load_state_dict hook drops the converted copy);
    in-place edits need invalidate(); release_reference_layout() frees the checkpoint-layout tensors and the layer still runs."""
    from oracle import svdq as O

    hT = torch.bfloat16
    la = O.make_synthetic_layer(256, 256, 32, fp4=fp4, hT=hT, seed=301)
    lb = O.make_synthetic_layer(256, 256, 32, fp4=fp4, hT=hT, seed=302)
    x = O.make_activations(200, 256, hT, seed=303).cuda().view(1, 200, 256)
    prec = "nvfp4" if fp4 else "int4"
    ma, _ = _module(la, prec)
    mb, sdb = _module(lb, prec)
    ya, yb = ma(x).clone(), mb(x).clone()
    assert O.rel_fro(ya.cpu(), yb.cpu()) > 0.5
    if fp4:
        ma.wtscale = lb.alpha
    ma.load_state_dict(sdb)
    assert torch.equal(ma(x), yb)
    # in-place edit of a parameter: invisible until invalidate()
    ma.bias.data.add_(1.0)
    assert torch.equal(ma(x), yb)
    ma.invalidate()
    assert O.rel_fro(ma(x).cpu(), (yb.float() + 1.0).cpu()) < 1e-2
    freed = mb.release_reference_layout()
    assert freed >= 256 * 128 and mb.qweight.numel() == 0
    assert torch.equal(mb(x), yb)


def test_raw_op_cache_invalidate_after_inplace_update():
    """the raw-tensor operator path (reference signature) caches converted copies per source tensor: `.data.copy_` does not bump
    `_version`, so the loader must call repack.invalidate() -- and then gets the new weights"""
    from nunchaku_b200 import repack
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda
    from nunchaku_b200.ops.quantize import svdq_quantize_w4a4_act_fuse_lora_cuda
    from gpu_util import ref_layout_params
    from oracle import svdq as O

    hT = torch.bfloat16
    la = O.make_synthetic_layer(256, 256, 32, fp4=False, hT=hT, seed=311)
    lb = O.make_synthetic_layer(256, 256, 32, fp4=False, hT=hT, seed=312)
    pa, pb = ref_layout_params(la), ref_layout_params(lb)
    x = O.make_activations(256, 256, hT, seed=313).cuda()

    def run(p):
        q, s, l = svdq_quantize_w4a4_act_fuse_lora_cuda(x, lora_down=p["proj_down"], smooth=p["smooth"])
        out = torch.empty(256, 256, dtype=hT, device="cuda")
        svdq_gemm_w4a4_cuda(act=q, wgt=p["qweight"], out=out, ascales=s, wscales=p["wscales"], lora_act_in=l, lora_up=p["proj_up"], bias=p["bias"])
        return out

    ya, yb = run(pa), run(pb)
    for k in ("qweight", "wscales", "bias", "smooth", "proj_down", "proj_up"):
        pa[k].data.copy_(pb[k])
    stale = run(pa)
    assert torch.equal(stale, ya)            # documented behaviour: stale until told
    for k in ("qweight", "wscales", "bias", "smooth", "proj_down", "proj_up"):
        repack.invalidate(pa[k])
    assert torch.equal(run(pa), yb)
    # a tensor object that moves to new storage evicts its old entries
    n0 = len(repack._cache)
    pa["qweight"].data = pb["qweight"].clone()
    run(pa)
    assert len(repack._cache) <= n0 + 1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_layer_on_second_gpu_while_first_is_current():
    """per-device launch state (max dynamic shared memory attribute, SM count, cluster occupancy): cuda:1 used while cuda:0 is current"""
    from oracle import svdq as O

    hT = torch.bfloat16
    for fp4 in (False, True):
        layer = O.make_synthetic_layer(512, 256, 32, fp4=fp4, hT=hT, seed=321)
        x = O.make_activations(300, 256, hT, seed=322)
        m0, _ = _module(layer, "nvfp4" if fp4 else "int4", device="cuda:0")
        y0 = m0(x.to("cuda:0").view(1, 300, 256))
        torch.cuda.set_device(0)
        m1, _ = _module(layer, "nvfp4" if fp4 else "int4", device="cuda:1")
        y1 = m1(x.to("cuda:1").view(1, 300, 256))
        torch.cuda.synchronize("cuda:1")
        assert torch.equal(y0.cpu(), y1.cpu())


@pytest.mark.parametrize("fp4", [False, True])
def test_sidecar_of_converted_weights(tmp_path, fp4):
    """save_b200 -> a fresh module adopts the file (no repack kernels, checkpoint-layout copy optional) -> bit-identical outputs;
    a side-car made from other parameters is refused"""
    from oracle import svdq as O

    hT = torch.bfloat16
    layer = O.make_synthetic_layer(256, 256, 32, fp4=fp4, hT=hT, seed=331)
    other = O.make_synthetic_layer(256, 256, 32, fp4=fp4, hT=hT, seed=332)
    prec = "nvfp4" if fp4 else "int4"
    m, sd = _module(layer, prec)
    x = O.make_activations(200, 256, hT, seed=333, smooth=layer.smooth).cuda().view(1, 200, 256)
    y = m(x)
    path = str(tmp_path / "l.b200")
    m.save_b200(path)
    # (a) same checkpoint loaded again: the side-car replaces the conversion
    m2, _ = _module(layer, prec)
    m2.load_b200(path)
    assert torch.equal(m2(x), y)
    # (b) side-car only: the module never sees the checkpoint-layout weight tensors
    from nunchaku_b200.models.linear import SVDQW4A4Linear

    m3 = SVDQW4A4Linear(256, 256, rank=32, bias=True, precision=prec, torch_dtype=hT, device="cuda")
    if fp4:
        m3.wtscale = layer.alpha
    m3.load_b200(path, require_source=False)
    assert m3.qweight.numel() == 0 and torch.equal(m3(x), y)
    # (c) a module holding OTHER parameters refuses it
    m4, _ = _module(other, prec)
    with pytest.raises(ValueError, match="different parameters"):
        m4.load_b200(path)
    torch.cuda.synchronize()
