"""Full-size parity for every shape and epilogue bench.py runs: too big for the CPU oracle in
seconds, so the SAME quantised operands are evaluated in fp64 with torch on the GPU and the epilogue chain is restated
there (size-independent property: the kernel is linear in the dequantised operands; the oracle pins the epilogue
arithmetic at small sizes in tests/test_gpu_fused.py and against the reference's kernels in tests/test_ref_gpu_golden.py).

  (4352, 3072 -> 9216)  RMSNorm + RoPE epilogue, 24 heads, real pack_rotemb table: row-major `out` AND out_q/out_k/out_v
  (4352, 3072 -> 12288) plain and + GELU
  (4352, 12288 -> 3072) auto dispatch (block_n = 0)
  (256,  3072 -> 12288) fused GELU -> quantise(next) + low-rank down (R = 32), then fc2 (256, 12288 -> 3072)
  (4096, 3072 -> 3072)  the primary shape, every tile configuration
INT4 and NVFP4, bf16 (+ fp16 on the primary shape).
"""
import pytest
import torch

import b200_layouts as L
from gpu_util import diag, ref_layout_params
from oracle import formats as F
from oracle import svdq as O
from test_gpu_gemm import _record

pytestmark = pytest.mark.gpu
DEV = "cuda"


class Case:
    """random quantised layer + random quantised activations, logical form on the CPU, dequantised fp64 on the GPU"""

    def __init__(self, M, K, N, fp4, hT, seed, R=32, unsigned=False):
        g = torch.Generator().manual_seed(seed)
        self.M, self.K, self.N, self.fp4, self.hT, self.R = M, K, N, fp4, hT, R
        Mp = (M + 255) // 256 * 256
        self.Mp = Mp
        self.alpha = 1.0
        self.wcscales = None
        if fp4:
            qw = torch.randint(0, 16, (N, K), generator=g, dtype=torch.int8)
            ws = torch.randint(40, 56, (N, K // 16), generator=g, dtype=torch.uint8)
            qa = torch.randint(0, 16, (Mp, K), generator=g, dtype=torch.int8)
            sa = torch.randint(40, 56, (K // 16, Mp), generator=g, dtype=torch.uint8)
            self.alpha = 0.37 / K ** 0.5
            self.wcscales = (1.0 + 0.1 * torch.randn(N, generator=g)).to(hT)
            self.a64 = (O.e2m1_decode(qa.to(DEV)).view(Mp, K // 16, 16) * O.e4m3_decode(sa.to(DEV)).t().unsqueeze(-1)).view(Mp, K)
            self.w64 = (O.e2m1_decode(qw.to(DEV)).view(N, K // 16, 16) * O.e4m3_decode(ws.to(DEV)).unsqueeze(-1)).view(N, K)
        else:
            qw = torch.randint(-8, 8, (N, K), generator=g, dtype=torch.int8)
            ws = ((torch.rand(N, K // 64, generator=g) * 0.5 + 0.75) / (4.6 * K ** 0.5)).to(hT)
            qa = torch.randint(0, 16, (Mp, K), generator=g, dtype=torch.int8) if unsigned else torch.randint(-8, 8, (Mp, K), generator=g, dtype=torch.int8)
            sa = (torch.rand(K // 64, Mp, generator=g) * 0.2 + 0.05).to(hT)
            self.a64 = (qa.to(DEV).double().view(Mp, K // 64, 64) * sa.to(DEV).double().t().unsqueeze(-1)).view(Mp, K)
            self.w64 = (qw.to(DEV).double().view(N, K // 64, 64) * ws.to(DEV).double().unsqueeze(-1)).view(N, K)
        self.unsigned = unsigned
        bias = (torch.randn(N, generator=g) * 0.1).to(hT)
        lu = (torch.randn(N, R, generator=g) * 0.05).to(hT)
        ld = (torch.randn(R, K, generator=g) * 0.05).to(hT)
        smooth = (torch.rand(K, generator=g) + 0.5).to(hT)
        self.la = torch.randn(Mp, R, generator=g)
        self.layer = O.SynthLayer(qw=qw, wscales=ws, bias=bias, smooth=smooth, lora_down=ld, lora_up=lu, wcscales=self.wcscales, alpha=self.alpha,
                                  fp4=fp4, hT=hT)
        self.params = ref_layout_params(self.layer)
        if fp4:
            self.act = L.pack_fp4(qa).to(DEV)
            self.asc = L.pack_sf_tiles(sa.t().contiguous()).view(torch.float8_e4m3fn).view(K // 16, Mp).to(DEV)
        else:
            self.act = L.pack_int4(qa, signed=not unsigned).to(DEV)
            self.asc = sa.to(DEV)

    def gemm_kwargs(self):
        p = self.params
        return dict(act=self.act, wgt=p["qweight"], ascales=self.asc, wscales=p["wscales"], lora_act_in=self.la.to(DEV), lora_up=p["proj_up"],
                    bias=p["bias"], fp4=self.fp4, alpha=self.alpha, wcscales=p["wcscales"], act_unsigned=self.unsigned)

    def pre_activation(self):
        """fp64 [Mp, N]: (A W^T) alpha wcscale + bias + hT(la) Lu^T"""
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            y = (self.a64 @ self.w64.t()) * self.alpha
            if self.wcscales is not None:
                y = y * self.wcscales.to(DEV).double().view(1, -1)
            y = y + self.layer.bias.to(DEV).double().view(1, -1)
            y = y + self.la.to(DEV).to(self.hT).double() @ self.layer.lora_up.to(DEV).double().t()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
        return y


def gelu64(x):
    return x * (0.5 + 0.5 * torch.tanh(0.79788456 * (x + 0.044715 * x ** 3)))


def run_gemm(case, block_n=0, **extra):
    from nunchaku_b200.ops import gemm as G

    G.BLOCK_N_OVERRIDE = block_n
    try:
        G.svdq_gemm_w4a4_cuda(**{**case.gemm_kwargs(), **extra})
        torch.cuda.synchronize()
    finally:
        G.BLOCK_N_OVERRIDE = 0


def check(name, got, want, tol):
    e = O.rel_fro(got.cpu(), want.cpu())
    _record("fullsize " + name, ours_vs_exact=e)
    assert not torch.isnan(got).any() and e <= tol, f"{name}: rel_fro {e:.3e}\n" + diag(name, got, want)


TOL = {torch.bfloat16: 3e-3, torch.float16: 8e-4}


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("block_n", [0, 128, 256, 512, 1024, 2048])
def test_primary_shape_every_tile_configuration(fp4, hT, block_n):
    if block_n >= 1024 and not fp4:
        pytest.skip("cluster kernel is NVFP4")
    if hT == torch.float16 and block_n not in (0, 2048):
        pytest.skip("fp16: auto + cluster")
    c = Case(4096, 3072, 3072, fp4, hT, seed=1)
    out = torch.full((c.M, c.N), float("nan"), dtype=hT, device=DEV)
    run_gemm(c, block_n, out=out)
    check(f"primary fp4={fp4} {hT} bn={block_n}", out, c.pre_activation()[: c.M], TOL[hT])


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("K,N,gelu", [(3072, 12288, False), (3072, 12288, True), (12288, 3072, False), (3072, 9216, False)])
def test_flux_single_stream_shapes_auto_dispatch(fp4, K, N, gelu):
    hT = torch.bfloat16
    c = Case(4352, K, N, fp4, hT, seed=K + N)
    out = torch.full((c.M, c.N), float("nan"), dtype=hT, device=DEV)
    run_gemm(c, 0, out=out, fuse_gelu=gelu)
    y = c.pre_activation()[: c.M]
    if gelu:
        y = gelu64(y.to(hT).double())
    check(f"flux {K}->{N} gelu={gelu} fp4={fp4}", out, y, TOL[hT] * (1.5 if gelu else 1.0))


def test_fc2_on_unsigned_activations_full_size():
    c = Case(4352, 12288, 3072, False, torch.bfloat16, seed=5, unsigned=True)
    out = torch.full((c.M, c.N), float("nan"), dtype=torch.bfloat16, device=DEV)
    run_gemm(c, 0, out=out)
    check("fc2 unsigned int4", out, c.pre_activation()[: c.M], TOL[torch.bfloat16])


def rope64(y, norm_q, norm_k, sin, cos, hT):
    """EpilogueRMSNormRope on hT-rounded pre-activations, fp64 (epilogues.cuh:269-425)"""
    M, N = y.shape
    y = y.to(hT).double()
    out = y.clone()
    H = N // 3 // 128
    for part, w in ((0, norm_q), (1, norm_k)):
        blk = y[:, part * (N // 3):(part + 1) * (N // 3)].view(M, H, 128)
        v = blk * torch.rsqrt((blk * blk).sum(-1, keepdim=True) / 128.0 + 1e-6) * w.double().view(1, 1, 128)
        x0, x1 = v[..., 0::2], v[..., 1::2]
        s, c = sin.double().view(M, 1, 64), cos.double().view(M, 1, 64)
        out[:, part * (N // 3):(part + 1) * (N // 3)] = torch.stack([x0 * c - x1 * s, x0 * s + x1 * c], dim=-1).view(M, -1)
    return out


@pytest.mark.parametrize("fp4", [False, True])
@pytest.mark.parametrize("block_n", [0, 128])
def test_qkv_rope_24_heads_full_size(fp4, block_n):
    hT = torch.bfloat16
    c = Case(4352, 3072, 9216, fp4, hT, seed=9)
    g = torch.Generator(device=DEV).manual_seed(10)
    norm_q = (1 + 0.2 * torch.randn(128, generator=g, device=DEV)).to(hT)
    norm_k = (1 + 0.2 * torch.randn(128, generator=g, device=DEV)).to(hT)
    ang = torch.rand(c.Mp, 64, generator=g, device=DEV) * 6.283
    sin, cos = torch.sin(ang).float(), torch.cos(ang).float()
    rot = F.pack_rotemb(sin, cos)
    out = torch.full((c.M, c.N), float("nan"), dtype=hT, device=DEV)
    run_gemm(c, block_n, out=out, norm_q=norm_q, norm_k=norm_k, rotary_emb=rot)
    want = rope64(c.pre_activation(), norm_q, norm_k, sin, cos, hT)
    check(f"qkv rope 24 heads fp4={fp4} bn={block_n}", out, want[: c.M], 4e-3)
    # EpiloguePackQKV at the same size: bit-identical to the row-major result, converted hT -> fp32 -> fp16, pad rows masked
    H = 24
    outs = [torch.full((1, H, c.Mp, 128), 7.0, dtype=torch.float16, device=DEV) for _ in range(3)]
    run_gemm(c, block_n, out_q=outs[0], out_k=outs[1], out_v=outs[2], attn_tokens=c.M, norm_q=norm_q, norm_k=norm_k, rotary_emb=rot)
    wantp = out.float().to(torch.float16).view(c.M, 3, H, 128).permute(1, 2, 0, 3)
    for part in range(3):
        got = outs[part][0]
        assert torch.equal(got[:, : c.M].contiguous().view(torch.int16), wantp[part].contiguous().view(torch.int16)), f"part {part}"
        if c.Mp > c.M:
            assert torch.isnan(got[:, c.M:]).all() if part == 1 else bool((got[:, c.M:] == 0).all())
    # the split route (plain GEMM into a scratch + csrc/rope.cu's RMSNorm / RoPE / pack kernel) writes the same bits
    outs2 = [torch.full((1, H, c.Mp, 128), 7.0, dtype=torch.float16, device=DEV) for _ in range(3)]
    scratch = torch.empty(c.Mp, c.N, dtype=hT, device=DEV)
    run_gemm(c, block_n, out_q=outs2[0], out_k=outs2[1], out_v=outs2[2], attn_tokens=c.M, norm_q=norm_q, norm_k=norm_k, rotary_emb=rot, qkv_scratch=scratch)
    for part in range(3):
        assert torch.equal(outs2[part].view(torch.int16), outs[part].view(torch.int16)), f"split route, part {part}"


@pytest.mark.parametrize("fp4", [False, True])
def test_text_stream_fused_mlp_full_size(fp4):
    """(256, 3072 -> 12288) fused GELU -> quantise(next) + low-rank down (R = 32) -> fc2 (256, 12288 -> 3072)"""
    hT = torch.bfloat16
    M, D, Hd = 256, 3072, 12288
    c1 = Case(M, D, Hd, fp4, hT, seed=21)
    c2 = Case(M, Hd, D, fp4, hT, seed=22, unsigned=not fp4)
    p2 = c2.params
    Mp = c1.Mp
    hid = torch.full((M, Hd), float("nan"), dtype=hT, device=DEV)
    qout = torch.zeros(Mp, Hd // 2, dtype=torch.uint8, device=DEV)
    osc = torch.zeros(Hd // 16, Mp, dtype=torch.float8_e4m3fn, device=DEV) if fp4 else torch.zeros(Hd // 64, Mp, dtype=hT, device=DEV)
    la2 = torch.full((Mp, 32), float("nan"), dtype=torch.float32, device=DEV)
    run_gemm(c1, 0, out=hid, qout=qout, oscales=osc, lora_down=p2["proj_down"], lora_act_out=la2, smooth_factor=p2["smooth"])
    y = gelu64(c1.pre_activation()[:M].to(hT).double())
    check(f"fused fc1 hidden fp4={fp4}", hid, y, 4.5e-3)
    # the hand-off tensors, re-derived by the oracle's quantiser from the tile the kernel itself stored
    g_h = hid.cpu()
    shift = 0.0 if fp4 else O.SHIFT_GELU
    ys = O.h_div(O.rn(g_h.double() + shift, hT), c2.layer.smooth.view(1, Hd))
    q_exp, s_exp = O._quantize_rows_fp4(ys) if fp4 else O._quantize_rows_int4(ys, unsigned=True)
    if fp4:
        codes = L.unpack_fp4(qout.cpu())[:M]
        scales = L.unpack_sf_tiles(osc.cpu().view(torch.uint8).reshape(-1), Mp, Hd // 16).t().contiguous()[:, :M]
        assert torch.equal(scales, s_exp)
        nz = (O.e4m3_decode(s_exp).t() != 0).repeat_interleave(16, dim=1)
        cmp = O.compare_codes(codes[nz], q_exp[nz], fp4=True)
    else:
        codes = L.unpack_int4(qout.cpu(), signed=False)[:M]
        assert torch.equal(osc.cpu()[:, :M].view(torch.int16), s_exp.view(torch.int16))
        cmp = O.compare_codes(codes, q_exp, fp4=False)
    assert cmp["frac"] <= 2e-3 and cmp["max_step"] <= 1, cmp
    e = O.rel_fro(la2.cpu()[:M], g_h.double() @ c2.layer.lora_down.double().t())
    assert e <= 1e-4, e
    # fc2 consuming exactly those tensors
    out2 = torch.full((M, D), float("nan"), dtype=hT, device=DEV)
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda

    svdq_gemm_w4a4_cuda(act=qout, wgt=p2["qweight"], out=out2, ascales=osc, wscales=p2["wscales"], lora_act_in=la2, lora_up=p2["proj_up"],
                        bias=p2["bias"], fp4=fp4, alpha=c2.alpha, wcscales=p2["wcscales"], act_unsigned=not fp4)
    torch.cuda.synchronize()
    if fp4:
        a64 = (O.e2m1_decode(L.unpack_fp4(qout.cpu()).to(DEV)).view(Mp, Hd // 16, 16) *
               O.e4m3_decode(L.unpack_sf_tiles(osc.cpu().view(torch.uint8).reshape(-1), Mp, Hd // 16).to(DEV)).unsqueeze(-1)).view(Mp, Hd)
    else:
        a64 = (L.unpack_int4(qout.cpu(), signed=False).to(DEV).double().view(Mp, Hd // 64, 64) * osc.double().t().unsqueeze(-1)).view(Mp, Hd)
    y2 = (a64 @ c2.w64.t()) * c2.alpha
    if c2.wcscales is not None:
        y2 = y2 * c2.wcscales.to(DEV).double().view(1, -1)
    y2 = y2 + c2.layer.bias.to(DEV).double().view(1, -1) + la2.to(hT).double() @ c2.layer.lora_up.to(DEV).double().t()
    check(f"fc2 after fused fc1 fp4={fp4}", out2, y2[:M], TOL[hT])
