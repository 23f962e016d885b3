"""GPU parity of the fused GEMM epilogues (through the C ABI) against the CPU oracle:
  * fc1 -> GELU -> [next layer's lora_down] -> 4-bit quantise for fc2   (launch_impl:282-310)
  * QKV projection -> per-head RMSNorm(Q, K) -> RoPE                      (launch_impl:347-406)

Tolerances:
  * `out` (the hT GELU / RoPE tile): same gates as test_gpu_gemm (1e-2 vs the reference-emulating
    oracle, 4e-3 / 1.5e-3 vs fp64 for bf16 / fp16; tanh.approx / rsqrt.approx add < 1e-3);
  * next-layer codes: re-derived by the ORACLE's quantiser from the kernel's own stored hT tile, so
    only tie flips of the approximate reciprocal remain: scales bit-exact, code mismatch fraction
    <= 2e-3, |step| <= 1;
  * lora_act_out vs fp64 matmul of the stored tile: rel-Frobenius <= 1e-4 (fp32 atomics order).
"""
import os

import pytest
import torch

import b200_layouts as L
from gpu_util import diag, ref_layout_params
from oracle import formats as F
from oracle import svdq as O
from test_gpu_gemm import _assert_parity, _record

pytestmark = pytest.mark.gpu


def _pack_act(qa, fp4, hT):
    Mp, K = qa.q.shape
    if fp4:
        return (L.pack_fp4(qa.q).cuda(),
                L.pack_sf_tiles(qa.scales.t().contiguous()).view(torch.float8_e4m3fn).view(K // 16, Mp).cuda())
    return L.pack_int4(qa.q, signed=True).cuda(), qa.scales.cuda()


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("with_out,R2", [(True, 32), (False, 16), (True, 0)])
def test_fused_gelu_quantize_next(fp4, hT, with_out, R2):
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda

    N, K, R, M = 384, 256, 32, 300
    layer = O.make_synthetic_layer(N, K, R, fp4=fp4, hT=hT, seed=51)
    g = torch.Generator().manual_seed(52)
    nxt_smooth = (torch.rand(N, generator=g) + 0.5).to(hT)
    nxt_ld = (torch.randn(max(R2, 16), N, generator=g) * 0.05).to(hT)[:R2] if R2 else None
    x = O.make_activations(M, K, hT, seed=53, smooth=layer.smooth)
    qa = O.quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down, fp4=fp4)
    Mp = qa.q.shape[0]
    params = ref_layout_params(layer)
    act, asc = _pack_act(qa, fp4, hT)
    out = torch.full((M, N), float("nan"), dtype=hT, device="cuda") if with_out else None
    qout = torch.zeros(Mp, N // 2, dtype=torch.uint8, device="cuda")
    osc = (torch.zeros(N // 16, Mp, dtype=torch.float8_e4m3fn, device="cuda") if fp4
           else torch.zeros(N // 64, Mp, dtype=hT, device="cuda"))
    la_out = torch.full((Mp, R2), float("nan"), dtype=torch.float32, device="cuda") if R2 else None
    svdq_gemm_w4a4_cuda(
        act=act, wgt=params["qweight"], out=out, qout=qout, ascales=asc, wscales=params["wscales"], oscales=osc,
        lora_act_in=qa.lora_act.cuda(), lora_up=params["proj_up"],
        lora_down=F.pack_lowrank(nxt_ld, down=True).cuda() if R2 else None, lora_act_out=la_out,
        bias=params["bias"], smooth_factor=F.pack_channel_vector(nxt_smooth).cuda(), fp4=fp4, alpha=layer.alpha,
        wcscales=params["wcscales"])
    torch.cuda.synchronize()

    kw = dict(qa=qa.q, ascales=qa.scales, qw=layer.qw, wscales=layer.wscales, hT=hT, M=M, bias=layer.bias,
              lora_act=qa.lora_act, lora_up=layer.lora_up, fp4=fp4, alpha=layer.alpha, wcscales=layer.wcscales,
              act="gelu", next_smooth=nxt_smooth, next_lora_down=nxt_ld)
    r_ref, r_ex = O.gemm_w4a4(mode="ref", **kw), O.gemm_w4a4(mode="exact", **kw)
    name = f"fused gelu+quant fp4={fp4} {hT} out={with_out} R2={R2}"
    if with_out:
        _assert_parity(out, r_ex.out, r_ref.out, hT, name)
        # re-derive the next layer's operands from the tile the kernel itself stored
        g_h = torch.zeros(Mp, N, dtype=hT)
        g_h[:M] = out.cpu()
        shift = 0.0 if fp4 else O.SHIFT_GELU
        ys = O.h_div(O.rn(g_h.double() + shift, hT), nxt_smooth.view(1, N))
        q_exp, s_exp = (O._quantize_rows_fp4(ys) if fp4 else O._quantize_rows_int4(ys, unsigned=True))
        if fp4:
            codes = L.unpack_fp4(qout.cpu())[:M]
            scales = L.unpack_sf_tiles(osc.cpu().view(torch.uint8).reshape(-1), Mp, N // 16).t().contiguous()[:, :M]
            assert torch.equal(scales, s_exp[:, :M]), diag("next scales", scales.float(), s_exp[:, :M].float())
            nz = (O.e4m3_decode(s_exp[:, :M]).t() != 0).repeat_interleave(16, dim=1)
            cmp = O.compare_codes(codes[nz], q_exp[:M][nz], fp4=True)
        else:
            codes = L.unpack_int4(qout.cpu(), signed=False)[:M]
            assert torch.equal(osc.cpu()[:, :M].view(torch.int16), s_exp[:, :M].view(torch.int16)), \
                diag("next scales", osc.float()[:, :M], s_exp[:, :M].float())
            cmp = O.compare_codes(codes, q_exp[:M], fp4=False)
        assert cmp["frac"] <= 2e-3 and cmp["max_step"] <= 1, (name, cmp)
        if R2:
            exp_la = g_h[:M].double() @ nxt_ld.double().t()
            e = O.rel_fro(la_out.cpu()[:M], exp_la)
            assert e <= 1e-4, diag("lora_act_out", la_out[:M], exp_la)
    else:
        # no stored tile: compare dequantised next-layer input and low-rank state with the oracle
        if fp4:
            deq = O.dequant(L.unpack_fp4(qout.cpu()), L.unpack_sf_tiles(osc.cpu().view(torch.uint8).reshape(-1), Mp, N // 16), True)[:M]
        else:
            deq = O.dequant(L.unpack_int4(qout.cpu(), signed=False), osc.cpu().t().contiguous(), False)[:M]
        deq_ref = O.dequant(r_ex.qout, r_ex.oscales.t().contiguous(), fp4)[:M]
        e = O.rel_fro(deq, deq_ref)
        _record(name + " deq", dtype=str(hT), ours_vs_exact=e)
        assert e <= 3e-2, diag("next act (dequantised)", deq, deq_ref)
        if R2:
            e2 = O.rel_fro(la_out.cpu()[:M], r_ex.lora_act_out[:M])
            assert e2 <= 1e-2, diag("lora_act_out", la_out[:M], r_ex.lora_act_out[:M])


@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("R2", [32, 16, 0])
def test_fused_gelu_quantize_next_int4_wide_tiles(hT, R2):
    """INT4 fused epilogue on 256-wide tiles (one TMEM accumulator, next-layer rank <= 32), two tiles per m-block so the
    down projection accumulates in TMEM across tiles before it is flushed: same outputs as the 128-wide kernel."""
    from nunchaku_b200.ops import gemm as G

    N, K, R, M = 512, 256, 32, 600
    layer = O.make_synthetic_layer(N, K, R, fp4=False, hT=hT, seed=151)
    g = torch.Generator().manual_seed(152)
    nxt_smooth = (torch.rand(N, generator=g) + 0.5).to(hT)
    nxt_ld = (torch.randn(32, N, generator=g) * 0.05).to(hT)[:R2] if R2 else None
    x = O.make_activations(M, K, hT, seed=153, smooth=layer.smooth)
    qa = O.quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down, fp4=False)
    Mp = qa.q.shape[0]
    params = ref_layout_params(layer)
    act, asc = _pack_act(qa, False, hT)
    res = {}
    for bn in (128, 256):
        out = torch.full((M, N), float("nan"), dtype=hT, device="cuda")
        qout = torch.zeros(Mp, N // 2, dtype=torch.uint8, device="cuda")
        osc = torch.zeros(N // 64, Mp, dtype=hT, device="cuda")
        la_out = torch.full((Mp, R2), float("nan"), dtype=torch.float32, device="cuda") if R2 else None
        G.BLOCK_N_OVERRIDE = bn
        try:
            G.svdq_gemm_w4a4_cuda(
                act=act, wgt=params["qweight"], out=out, qout=qout, ascales=asc, wscales=params["wscales"], oscales=osc,
                lora_act_in=qa.lora_act.cuda(), lora_up=params["proj_up"],
                lora_down=F.pack_lowrank(nxt_ld, down=True).cuda() if R2 else None, lora_act_out=la_out, bias=params["bias"],
                smooth_factor=F.pack_channel_vector(nxt_smooth).cuda(), fp4=False)
        finally:
            G.BLOCK_N_OVERRIDE = 0
        torch.cuda.synchronize()
        res[bn] = (out.cpu(), qout.cpu(), osc.cpu(), None if la_out is None else la_out.cpu())
    a, b = res[128], res[256]
    assert torch.equal(a[0].view(torch.int16), b[0].view(torch.int16))     # the hT GELU tile
    assert torch.equal(a[1], b[1]) and torch.equal(a[2].view(torch.int16), b[2].view(torch.int16))   # codes + scales
    if R2:
        assert (a[3][:M] - b[3][:M]).norm() <= 1e-5 * a[3][:M].norm()        # fp32 summation order only
        ref = (a[0][:M].double() @ nxt_ld.double().t())
        assert (b[3][:M].double() - ref).norm() <= 1e-4 * ref.norm()


@pytest.mark.parametrize("precision", ["int4", "nvfp4"])
def test_fused_gelu_mlp_module_chain(precision):
    """fused_gelu_mlp(x, fc1, fc2) (ops/fused.py:14-79) end to end vs the oracle chain."""
    from nunchaku_b200.models.linear import SVDQW4A4Linear
    from nunchaku_b200.ops.fused import fused_gelu_mlp

    fp4 = precision == "nvfp4"
    hT = torch.bfloat16
    D, H, R, M = 256, 512, 32, 200
    l1 = O.make_synthetic_layer(H, D, R, fp4=fp4, hT=hT, seed=61)
    l2 = O.make_synthetic_layer(D, H, R, fp4=fp4, hT=hT, seed=62)
    x = O.make_activations(M, D, hT, seed=63, smooth=l1.smooth)

    def mk(layer, K, N, unsigned):
        p = ref_layout_params(layer)
        m = SVDQW4A4Linear(K, N, rank=R, bias=True, precision=precision, act_unsigned=unsigned, torch_dtype=hT, device="cuda")
        sd = {"qweight": p["qweight"], "wscales": p["wscales"], "bias": p["bias"], "smooth_factor": p["smooth"],
              "smooth_factor_orig": p["smooth"], "proj_down": p["proj_down"], "proj_up": p["proj_up"]}
        if fp4:
            sd["wcscales"] = p["wcscales"]
            m.wtscale = layer.alpha
        m.load_state_dict(sd)
        return m

    fc1, fc2 = mk(l1, D, H, False), mk(l2, H, D, not fp4)
    y = fused_gelu_mlp(x.cuda().view(1, M, D), fc1, fc2).view(M, D)
    torch.cuda.synchronize()
    qa = O.quantize_w4a4_act_fuse_lora(x, l1.smooth, l1.lora_down, fp4=fp4)
    outs = {}
    for mode in ("ref", "exact"):
        r1 = O.gemm_w4a4(qa=qa.q, ascales=qa.scales, qw=l1.qw, wscales=l1.wscales, hT=hT, M=M, bias=l1.bias,
                         lora_act=qa.lora_act, lora_up=l1.lora_up, fp4=fp4, alpha=l1.alpha, wcscales=l1.wcscales,
                         act="gelu", next_smooth=l2.smooth, next_lora_down=l2.lora_down, want_out=False, mode=mode)
        outs[mode] = O.gemm_w4a4(qa=r1.qout, ascales=r1.oscales, qw=l2.qw, wscales=l2.wscales, hT=hT, M=M, bias=l2.bias,
                                 lora_act=r1.lora_act_out, lora_up=l2.lora_up, fp4=fp4, alpha=l2.alpha,
                                 wcscales=l2.wcscales, mode=mode).out
    e_ref, e_ex = O.rel_fro(y.cpu(), outs["ref"]), O.rel_fro(y.cpu(), outs["exact"])
    _record(f"fused_gelu_mlp {precision}", dtype=str(hT), ours_vs_exact=e_ex, ours_vs_ref=e_ref,
            ref_vs_exact=O.rel_fro(outs["ref"], outs["exact"]))
    # two quantisers in the chain: tie flips of the intermediate 4-bit codes dominate (see test_gpu_gemm)
    assert e_ref <= 3e-2 and e_ex <= 3e-2, f"{precision}: vs ref {e_ref:.3e} vs exact {e_ex:.3e}\n" + diag("mlp", y, outs["exact"])


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
def test_qkv_rmsnorm_rope(fp4, hT):
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda

    H, K, R, M = 2, 256, 32, 300
    N = 3 * H * 128
    layer = O.make_synthetic_layer(N, K, R, fp4=fp4, hT=hT, seed=71)
    g = torch.Generator().manual_seed(72)
    norm_q = (1.0 + 0.2 * torch.randn(128, generator=g)).to(hT)
    norm_k = (1.0 + 0.2 * torch.randn(128, generator=g)).to(hT)
    x = O.make_activations(M, K, hT, seed=73, smooth=layer.smooth)
    qa = O.quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down, fp4=fp4)
    Mp = qa.q.shape[0]
    ang = torch.rand(Mp, 64, generator=g) * 6.283
    rsin, rcos = torch.sin(ang), torch.cos(ang)
    params = ref_layout_params(layer)
    act, asc = _pack_act(qa, fp4, hT)
    out = torch.full((M, N), float("nan"), dtype=hT, device="cuda")
    svdq_gemm_w4a4_cuda(act=act, wgt=params["qweight"], out=out, ascales=asc, wscales=params["wscales"],
                        lora_act_in=qa.lora_act.cuda(), lora_up=params["proj_up"], bias=params["bias"], fp4=fp4,
                        alpha=layer.alpha, wcscales=params["wcscales"], norm_q=norm_q.cuda(), norm_k=norm_k.cuda(),
                        rotary_emb=F.pack_rotemb(rsin, rcos).cuda())
    torch.cuda.synchronize()
    kw = dict(qa=qa.q, ascales=qa.scales, qw=layer.qw, wscales=layer.wscales, hT=hT, M=M, bias=layer.bias,
              lora_act=qa.lora_act, lora_up=layer.lora_up, fp4=fp4, alpha=layer.alpha, wcscales=layer.wcscales,
              rope=(norm_q, norm_k, rsin, rcos))
    y_ref, y_ex = O.gemm_w4a4(mode="ref", **kw).out, O.gemm_w4a4(mode="exact", **kw).out
    _assert_parity(out, y_ex, y_ref, hT, f"qkv rmsnorm+rope fp4={fp4} {hT}")
    # V third untouched by norm/rope: equals the plain epilogue
    kw.pop("rope")
    v_ex = O.gemm_w4a4(mode="exact", **kw).out[:, 2 * N // 3:]
    assert O.rel_fro(out.cpu()[:, 2 * N // 3:], v_ex) <= (4e-3 if hT == torch.bfloat16 else 1.5e-3)


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
def test_qkv_pack_outputs(fp4, hT):
    """EpiloguePackQKV (epilogues.cuh:427-550): the RMSNorm+RoPE result delivered as three fp16 [1, H, Tpad, 128]
    tensors.  Bit-exact against the same kernel's row-major `out` (converted hT -> fp32 -> fp16 like
    convert_half2); pad rows (>= attn_tokens) are 0 for Q / V and NaN for K; memory outside the views is untouched.
    The element order inside a head is plain row-major (ours), not the reference attention kernel's fragment order."""
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda

    H, K, R, M = 2, 256, 32, 300
    N = 3 * H * 128
    layer = O.make_synthetic_layer(N, K, R, fp4=fp4, hT=hT, seed=81)
    g = torch.Generator().manual_seed(82)
    norm_q = (1.0 + 0.2 * torch.randn(128, generator=g)).to(hT)
    norm_k = (1.0 + 0.2 * torch.randn(128, generator=g)).to(hT)
    x = O.make_activations(M, K, hT, seed=83, smooth=layer.smooth)
    qa = O.quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down, fp4=fp4)
    Mp = qa.q.shape[0]
    ang = torch.rand(Mp, 64, generator=g) * 6.283
    rot = F.pack_rotemb(torch.sin(ang), torch.cos(ang)).cuda()
    params = ref_layout_params(layer)
    act, asc = _pack_act(qa, fp4, hT)
    common = dict(act=act, wgt=params["qweight"], ascales=asc, wscales=params["wscales"], lora_act_in=qa.lora_act.cuda(),
                  lora_up=params["proj_up"], bias=params["bias"], fp4=fp4, alpha=layer.alpha, wcscales=params["wcscales"],
                  norm_q=norm_q.cuda(), norm_k=norm_k.cuda(), rotary_emb=rot)
    out = torch.empty(Mp, N, dtype=hT, device="cuda")
    svdq_gemm_w4a4_cuda(out=out, **common)
    # joint-block style: this stream's tokens live at rows [256, 256 + Mp) of a longer sequence
    pad0 = 256
    full = [torch.full((1, H, pad0 + Mp, 128), 7.0, dtype=torch.float16, device="cuda") for _ in range(3)]
    views = [t[:, :, pad0:] for t in full]
    svdq_gemm_w4a4_cuda(out_q=views[0], out_k=views[1], out_v=views[2], attn_tokens=M, **common)
    torch.cuda.synchronize()
    want = out.float().to(torch.float16).view(Mp, 3, H, 128).permute(1, 2, 0, 3).cpu()  # [3, H, Mp, 128]
    for part, (t, v) in enumerate(zip(full, views)):
        got = v[0].cpu()
        assert torch.equal(got[:, :M].view(torch.int16), want[part][:, :M].view(torch.int16)), f"part {part}"
        if part == 1:
            assert torch.isnan(got[:, M:]).all()
        else:
            assert (got[:, M:] == 0).all()
        assert (t[:, :, :pad0] == 7.0).all()
    with pytest.raises(ValueError):
        svdq_gemm_w4a4_cuda(out_q=views[0], out_k=views[1], out_v=None, attn_tokens=M, **common)


@pytest.mark.parametrize("precision", ["nvfp4", "int4"])
def test_fused_gelu_mlp_both_routes_agree(precision):
    """MLP: fused fc1 epilogue vs plain GEMM + GELU followed by the activation quantizer (ops.fused.FUSE_FC1_EPILOGUE; INT4 uses
    the quantizer's shift_unsigned mode).  Same arithmetic on both routes: the 4-bit codes and scales handed to fc2 must be
    bit-identical; only the fp32 summation order of fc2's low-rank hidden state differs."""
    from nunchaku_b200.models.linear import SVDQW4A4Linear
    from nunchaku_b200.ops import fused as FU
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda
    from nunchaku_b200.ops.quantize import svdq_quantize_w4a4_act_fuse_lora_cuda

    fp4 = precision == "nvfp4"
    hT = torch.bfloat16
    D, H, R, M = 256, 512, 32, 300
    l1 = O.make_synthetic_layer(H, D, R, fp4=fp4, hT=hT, seed=161)
    l2 = O.make_synthetic_layer(D, H, R, fp4=fp4, hT=hT, seed=162)
    x = O.make_activations(M, D, hT, seed=163, smooth=l1.smooth)

    def mk(layer, K, N, unsigned):
        p = ref_layout_params(layer)
        m = SVDQW4A4Linear(K, N, rank=R, bias=True, precision=precision, act_unsigned=unsigned, torch_dtype=hT, device="cuda")
        sd = {"qweight": p["qweight"], "wscales": p["wscales"], "bias": p["bias"], "smooth_factor": p["smooth"],
              "smooth_factor_orig": p["smooth"], "proj_down": p["proj_down"], "proj_up": p["proj_up"]}
        if fp4:
            sd["wcscales"] = p["wcscales"]
            m.wtscale = layer.alpha
        m.load_state_dict(sd)
        return m

    fc1, fc2 = mk(l1, D, H, False), mk(l2, H, D, not fp4)
    xd = x.cuda()
    # the hand-off tensors of both routes
    qx, asc, la = fc1.quantize(xd)
    Mp = qx.shape[0]
    q_f = torch.zeros(Mp, H // 2, dtype=torch.uint8, device="cuda")
    s_f = torch.zeros(H // 16, Mp, dtype=torch.float8_e4m3fn, device="cuda") if fp4 else torch.zeros(H // 64, Mp, dtype=hT, device="cuda")
    la_f = torch.zeros(Mp, R, dtype=torch.float32, device="cuda")
    common = dict(act=qx, wgt=fc1.qweight, ascales=asc, wscales=fc1.wscales, lora_act_in=la, lora_up=fc1.proj_up, bias=fc1.bias, fp4=fp4,
                  alpha=fc1.wtscale, wcscales=fc1.wcscales)
    svdq_gemm_w4a4_cuda(qout=q_f, oscales=s_f, lora_down=fc2.proj_down, lora_act_out=la_f, smooth_factor=fc2.smooth_factor, **common)
    hid = torch.empty(M, H, dtype=hT, device="cuda")
    svdq_gemm_w4a4_cuda(out=hid, fuse_gelu=True, **common)
    q_s, s_s, la_s = svdq_quantize_w4a4_act_fuse_lora_cuda(hid, lora_down=fc2.proj_down, smooth=fc2.smooth_factor, fp4=fp4, shift_unsigned=not fp4)
    torch.cuda.synchronize()
    assert torch.equal(q_f[:M].cpu(), q_s[:M].cpu()), "4-bit codes handed to fc2 differ between the routes"
    if fp4:
        # scale tiles: only the rows < M are defined
        assert torch.equal(L.unpack_sf_tiles(s_f.view(torch.uint8).cpu().reshape(-1), Mp, H // 16)[:M], L.unpack_sf_tiles(s_s.view(torch.uint8).cpu().reshape(-1), Mp, H // 16)[:M])
    else:
        assert torch.equal(s_f[:, :M].cpu().view(torch.int16), s_s[:, :M].cpu().view(torch.int16))
    assert (la_f[:M] - la_s[:M]).norm() <= 1e-4 * la_s[:M].norm()
    ys = {}
    for route in (True, False):
        FU.FUSE_FC1_EPILOGUE = route
        try:
            ys[route] = FU.fused_gelu_mlp(xd.view(1, M, D), fc1, fc2).view(M, D).float().cpu()
        finally:
            FU.FUSE_FC1_EPILOGUE = None
    assert (ys[True] - ys[False]).norm() <= 2e-3 * ys[True].norm()


@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("block_n,H", [(1024, 2), (2048, 4), (1024, 3)])
def test_qkv_rmsnorm_rope_cluster_kernel(hT, block_n, H):
    """NVFP4 cluster kernel's RMSNorm + RoPE epilogue (one 128-wide head per epilogue group; H = 3: a 256-wide tile straddles
    the Q/K boundary) and its PackQKV variant, against the oracle and against its own row-major result."""
    from nunchaku_b200.ops import gemm as G
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda

    if (3 * H * 128) % (256 * (block_n // 1024)) != 0:
        pytest.skip("N must be a multiple of the cluster tile width")
    K, R, M = 256, 32, 300
    N = 3 * H * 128
    layer = O.make_synthetic_layer(N, K, R, fp4=True, hT=hT, seed=271)
    g = torch.Generator().manual_seed(272)
    norm_q = (1.0 + 0.2 * torch.randn(128, generator=g)).to(hT)
    norm_k = (1.0 + 0.2 * torch.randn(128, generator=g)).to(hT)
    x = O.make_activations(M, K, hT, seed=273, smooth=layer.smooth)
    qa = O.quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down, fp4=True)
    Mp = qa.q.shape[0]
    ang = torch.rand(Mp, 64, generator=g) * 6.283
    rsin, rcos = torch.sin(ang), torch.cos(ang)
    params = ref_layout_params(layer)
    act, asc = _pack_act(qa, True, hT)
    common = dict(act=act, wgt=params["qweight"], ascales=asc, wscales=params["wscales"], lora_act_in=qa.lora_act.cuda(),
                  lora_up=params["proj_up"], bias=params["bias"], fp4=True, alpha=layer.alpha, wcscales=params["wcscales"],
                  norm_q=norm_q.cuda(), norm_k=norm_k.cuda(), rotary_emb=F.pack_rotemb(rsin, rcos).cuda())
    out = torch.full((Mp, N), float("nan"), dtype=hT, device="cuda")
    outs = [torch.full((1, H, Mp, 128), 7.0, dtype=torch.float16, device="cuda") for _ in range(3)]
    G.BLOCK_N_OVERRIDE = block_n
    try:
        svdq_gemm_w4a4_cuda(out=out, **common)          # default route: plain GEMM + in-place RMSNorm/RoPE kernel (csrc/rope.cu)
        out_fused = torch.full((Mp, N), float("nan"), dtype=hT, device="cuda")
        os.environ["NB200_ROPE_SPLIT"] = "0"            # the same call with the epilogue fused
        try:
            svdq_gemm_w4a4_cuda(out=out_fused, **common)
        finally:
            del os.environ["NB200_ROPE_SPLIT"]
        svdq_gemm_w4a4_cuda(out_q=outs[0], out_k=outs[1], out_v=outs[2], attn_tokens=M, **common)
        torch.cuda.synchronize()
    finally:
        G.BLOCK_N_OVERRIDE = 0
    assert torch.equal(out.view(torch.int16), out_fused.view(torch.int16)), "split and fused RoPE routes must agree bit for bit"
    kw = dict(qa=qa.q, ascales=qa.scales, qw=layer.qw, wscales=layer.wscales, hT=hT, M=M, bias=layer.bias, lora_act=qa.lora_act,
              lora_up=layer.lora_up, fp4=True, alpha=layer.alpha, wcscales=layer.wcscales, rope=(norm_q, norm_k, rsin, rcos))
    _assert_parity(out[:M], O.gemm_w4a4(mode="exact", **kw).out, O.gemm_w4a4(mode="ref", **kw).out, hT, f"cluster rope bn={block_n} H={H} {hT}")
    want = out.float().to(torch.float16).view(Mp, 3, H, 128).permute(1, 2, 0, 3).cpu()
    for part in range(3):
        got = outs[part][0].cpu()
        assert torch.equal(got[:, :M].view(torch.int16), want[part][:, :M].view(torch.int16)), f"part {part}"
        assert torch.isnan(got[:, M:]).all() if part == 1 else bool((got[:, M:] == 0).all())


@pytest.mark.parametrize("precision", ["nvfp4", "int4"])
@pytest.mark.parametrize("M,D,H", [(300, 256, 512), (256, 3072, 12288)])
def test_fused_handoff_is_bit_stable(precision, M, D, H):
    """The fused fc1 epilogue's low-rank hidden state for fc2 is a sum over CTAs (several CTAs share a 128-row block, one CTA may span two): with
    the reduction workspace the partials are added in a fixed order by the last CTA to arrive, so repeated launches give identical bits
    (the reference adds them with red.global.add.f32) -- and they agree with the atomics path to fp32 rounding."""
    from nunchaku_b200.models.linear import SVDQW4A4Linear
    from nunchaku_b200.ops import gemm as G

    hT = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(M + D)

    def mk(K, N, unsigned):
        m = SVDQW4A4Linear(K, N, rank=32, bias=True, precision=precision, act_unsigned=unsigned, torch_dtype=hT, device="cuda")
        with torch.no_grad():
            m.qweight.copy_(torch.randint(-128, 128, m.qweight.shape, generator=g, device="cuda", dtype=torch.int8))
            if precision == "nvfp4":
                m.wscales.copy_(torch.randint(0x30, 0x38, m.wscales.shape, generator=g, device="cuda", dtype=torch.uint8).view(torch.float8_e4m3fn))
                m.wtscale = 1.0 / (2.6 * 0.72 * K ** 0.5)
                m.wcscales.copy_((1.0 + 0.05 * torch.randn(N, generator=g, device="cuda")).to(hT))
            else:
                m.wscales.copy_(((0.75 + 0.5 * torch.rand(m.wscales.shape, generator=g, device="cuda")) / (4.6 * K ** 0.5)).to(hT))
            m.bias.copy_((0.1 * torch.randn(N, generator=g, device="cuda")).to(hT))
            m.smooth_factor.copy_((0.75 + 0.5 * torch.rand(K, generator=g, device="cuda")).to(hT))
            m.proj_down.copy_((torch.randn(K, 32, generator=g, device="cuda") / K ** 0.5).to(hT))
            m.proj_up.copy_((0.1 * torch.randn(N, 32, generator=g, device="cuda") / 32 ** 0.5).to(hT))
        m.invalidate()
        return m

    fc1, fc2 = mk(D, H, False), mk(H, D, precision == "int4")
    x = torch.randn(M, D, generator=g, device="cuda").to(hT)
    q, s, la = fc1.quantize(x)
    runs = [fc1.quantize_next(q, s, la, fc2) for _ in range(4)]
    torch.cuda.synchronize()
    for q2, s2, la2 in runs[1:]:
        assert torch.equal(la2.view(torch.int32), runs[0][2].view(torch.int32))
        assert torch.equal(q2, runs[0][0]) and torch.equal(s2.view(torch.uint8), runs[0][1].view(torch.uint8))
    # the atomics path (no workspace) sums the same partials in a different order
    orig = G._reduce_workspace
    G._reduce_workspace = lambda Mp, rank, device: torch.zeros(0, dtype=torch.uint8, device=device)
    try:
        la_atomic = fc1.quantize_next(q, s, la, fc2)[2]
    finally:
        G._reduce_workspace = orig
    torch.cuda.synchronize()
    assert O.rel_fro(la_atomic.cpu(), runs[0][2].cpu()) <= 1e-6
