"""Generate GPU golden vectors by running the UNMODIFIED reference kernels on the B200.

The reference has no CPU implementation and no operator-level golden vectors for this path, and its
own build refuses sm_100 (setup.py:54).  Its kernel translation units do compile for sm_100a
(oracle/ref_build/build_ref.sh -> oracle/_ref/libnunchaku_ref.so; INT4 only -- the NVFP4 kernels need sm_120a's
``mma.sync ... block_scale`` and compile to a trap), so the oracle is pinned against what the
reference's kernels produce on this very GPU:

    gpurun -- python tests/golden/make_ref_gpu_golden.py        # writes gpurun_out/ref_gpu_golden.npz (+ report)
    cp gpurun_out/ref_gpu_golden.npz tests/golden/              # committed fixture

Every case stores the LOGICAL inputs (so tests/test_ref_gpu_golden.py can feed the CPU oracle on any machine)
and the RAW bytes the reference wrote (so the layout formulas in oracle/formats.py are part of what is pinned).
The script also prints the oracle-vs-reference comparison immediately (gpurun_out/ref_gpu_golden_report.json).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import formats as F  # noqa: E402
from oracle import refgpu as R  # noqa: E402
from oracle import svdq as O  # noqa: E402

OUT_DIR = os.path.join(ROOT, "gpurun_out")
DEV = "cuda"


def npy(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.view(torch.int16).numpy().copy()
    if t.dtype == torch.float8_e4m3fn:
        return t.view(torch.uint8).numpy().copy()
    return t.numpy().copy()


def layer_to_ref(layer: O.SynthLayer):
    """logical synthetic layer -> the reference's packed checkpoint tensors on the GPU"""
    p = {
        "qweight": F.pack_qweight(layer.qw).to(DEV),
        "wscales": F.pack_group_scales(layer.wscales).to(DEV),
        "bias": F.pack_channel_vector(layer.bias).to(DEV),
        "smooth": F.pack_channel_vector(layer.smooth).to(DEV),
        "lora_up": F.pack_lowrank(layer.lora_up, down=False).to(DEV),
        "lora_down": F.pack_lowrank(layer.lora_down, down=True).to(DEV),
    }
    return p


def store_layer(out: dict, pre: str, layer: O.SynthLayer):
    out[pre + "qw"] = npy(layer.qw)
    out[pre + "wscales"] = npy(layer.wscales)
    out[pre + "bias"] = npy(layer.bias)
    out[pre + "smooth"] = npy(layer.smooth)
    out[pre + "lora_up"] = npy(layer.lora_up)
    out[pre + "lora_down"] = npy(layer.lora_down)


def run_case(tag: str, hT, N, K, Rk, M, out: dict, report: dict, seed: int, *, gelu_quant_next: int = 0, rope: bool = False,
             silu: bool = False, lora_scales=None, glu: bool = False):
    pre = f"{tag}."
    layer = O.make_synthetic_layer(N, K, Rk, fp4=False, hT=hT, seed=seed)
    x = O.make_activations(M, K * (2 if glu else 1), hT=hT, seed=seed + 100, smooth=None if glu else layer.smooth)
    p = layer_to_ref(layer)
    store_layer(out, pre, layer)
    out[pre + "x"] = npy(x)
    meta = dict(N=N, K=K, R=Rk, M=M, dtype=str(hT).split(".")[-1], glu=glu, silu=silu, rope=rope, gelu_quant_next=gelu_quant_next,
                lora_scales=lora_scales)

    # ---- reference quantizer ----
    act, asc, la = R.quantize_w4a4_act_fuse_lora(x.to(DEV), p["lora_down"], p["smooth"], fuse_glu=glu)
    torch.cuda.synchronize()
    out[pre + "ref_act"] = npy(act)
    out[pre + "ref_ascales"] = npy(asc)
    out[pre + "ref_lora_act"] = npy(la)
    Mp = act.shape[0]
    # oracle on the same inputs
    qa = O.quantize_w4a4_act_fuse_lora(x, layer.smooth, layer.lora_down, fuse_glu=glu)
    codes = F.unpack_ref_act(act.cpu(), signed=True)
    scales = F.unpack_ref_ascales(asc.cpu()).t().contiguous()          # [G, Mp]
    lact = F.unpack_ref_lora_act(la.cpu())
    rep = {"meta": meta}
    rep["quant_codes"] = O.compare_codes(qa.q, codes)
    rep["quant_scales_equal"] = bool(torch.equal(qa.scales.view(torch.int16), scales.view(torch.int16)))
    rep["quant_lora_act_rel"] = O.rel_fro(qa.lora_act, lact)

    # ---- reference GEMM on the reference's own quantized activations ----
    kw = dict(ascales=asc, wscales=p["wscales"], lora_act_in=la, lora_up=p["lora_up"], bias=p["bias"], lora_scales=lora_scales,
              fuse_silu=silu)
    okw = dict(qa=codes, ascales=scales, qw=layer.qw, wscales=layer.wscales, hT=hT, M=M,
               bias=layer.bias, lora_act=lact, lora_up=layer.lora_up, lora_scales=lora_scales, act="silu" if silu else "none")
    if rope:
        g = torch.Generator().manual_seed(seed + 7)
        norm_q = (1.0 + 0.1 * torch.randn(128, generator=g)).to(hT)
        norm_k = (1.0 + 0.1 * torch.randn(128, generator=g)).to(hT)
        theta = torch.rand(Mp, 64, generator=g) * 6.28318
        rsin, rcos = torch.sin(theta).float(), torch.cos(theta).float()
        rot = F.pack_rotemb(rsin, rcos).view(1, Mp, 128).contiguous().to(DEV)
        out[pre + "norm_q"], out[pre + "norm_k"], out[pre + "rope_sin"], out[pre + "rope_cos"] = npy(norm_q), npy(norm_k), npy(rsin), npy(rcos)
        y = torch.empty(M, N, dtype=hT, device=DEV)
        R.gemm_w4a4(act, p["qweight"], out=y, norm_q=norm_q.to(DEV), norm_k=norm_k.to(DEV), rotary_emb=rot, **kw)
        torch.cuda.synchronize()
        out[pre + "ref_out"] = npy(y)
        for mode in ("ref", "exact"):
            o = O.gemm_w4a4(rope=(norm_q, norm_k, rsin[:Mp], rcos[:Mp]), mode=mode, **okw).out
            rep[f"gemm_rope_vs_{mode}"] = O.rel_fro(y.cpu(), o)
        if hT == torch.float16:   # EpiloguePackQKV: the reference attention kernel's private layout, raw
            H = N // 384
            oq = torch.zeros(1, H, Mp, 128, dtype=torch.float16, device=DEV)
            ok_ = torch.zeros_like(oq)
            ov = torch.zeros_like(oq)
            R.gemm_w4a4(act, p["qweight"], out=y, norm_q=norm_q.to(DEV), norm_k=norm_k.to(DEV), rotary_emb=rot, out_q=oq, out_k=ok_, out_v=ov,
                        attn_tokens=M, **kw)
            torch.cuda.synchronize()
            out[pre + "ref_out_q"], out[pre + "ref_out_k"], out[pre + "ref_out_v"] = npy(oq), npy(ok_), npy(ov)
    elif gelu_quant_next:
        nxt = O.make_synthetic_layer(gelu_quant_next, N, Rk, fp4=False, hT=hT, seed=seed + 1)
        pn = layer_to_ref(nxt)
        store_layer(out, pre + "next.", nxt)
        qout = torch.empty(Mp, N // 2, dtype=torch.uint8, device=DEV)
        osc = torch.empty(N // 64, Mp, dtype=hT, device=DEV)
        lao = torch.empty(Mp, Rk, dtype=torch.float32, device=DEV)
        R.gemm_w4a4(act, p["qweight"], qout=qout, oscales=osc, lora_down=pn["lora_down"], lora_act_out=lao, smooth_factor=pn["smooth"], **kw)
        torch.cuda.synchronize()
        out[pre + "ref_qout"], out[pre + "ref_oscales"], out[pre + "ref_lora_act_out"] = npy(qout), npy(osc), npy(lao)
        res = O.gemm_w4a4(act="gelu", next_smooth=nxt.smooth, next_lora_down=nxt.lora_down, want_out=False,
                          **{k: v for k, v in okw.items() if k != "act"})
        qc = F.unpack_ref_act(qout.cpu(), signed=False)
        qs = F.unpack_ref_ascales(osc.cpu()).t().contiguous()
        rep["fused_codes"] = O.compare_codes(res.qout, qc)
        rep["fused_scales_mismatch_frac"] = float((res.oscales.view(torch.int16) != qs.view(torch.int16)).double().mean())
        rep["fused_lora_act_out_rel"] = O.rel_fro(res.lora_act_out, F.unpack_ref_lora_act(lao.cpu()))
        # fc2 consumes the unsigned activations (act_unsigned = true, Linear.cpp:294)
        y2 = torch.empty(M, gelu_quant_next, dtype=hT, device=DEV)
        R.gemm_w4a4(qout, pn["qweight"], out=y2, ascales=osc, wscales=pn["wscales"], lora_act_in=lao, lora_up=pn["lora_up"], bias=pn["bias"],
                    act_unsigned=True)
        torch.cuda.synchronize()
        out[pre + "ref_out2"] = npy(y2)
        for mode in ("ref", "exact"):
            o2 = O.gemm_w4a4(qa=qc, ascales=qs, qw=nxt.qw, wscales=nxt.wscales, hT=hT, M=M, bias=nxt.bias, lora_act=F.unpack_ref_lora_act(lao.cpu()),
                             lora_up=nxt.lora_up, mode=mode).out
            rep[f"gemm_unsigned_vs_{mode}"] = O.rel_fro(y2.cpu(), o2)
    else:
        y = torch.empty(M, N, dtype=hT, device=DEV)
        R.gemm_w4a4(act, p["qweight"], out=y, **kw)
        torch.cuda.synchronize()
        out[pre + "ref_out"] = npy(y)
        # run-to-run noise of the reference itself (the quantizer's fp32 atomics)
        act2, asc2, la2 = R.quantize_w4a4_act_fuse_lora(x.to(DEV), p["lora_down"], p["smooth"], fuse_glu=glu)
        y_b = torch.empty_like(y)
        R.gemm_w4a4(act2, p["qweight"], out=y_b, **{**kw, "ascales": asc2, "lora_act_in": la2})
        torch.cuda.synchronize()
        rep["ref_run_to_run"] = O.rel_fro(y_b.cpu(), y.cpu())
        for mode in ("ref", "exact"):
            o = O.gemm_w4a4(mode=mode, **okw).out
            rep[f"gemm_vs_{mode}"] = O.rel_fro(y.cpu(), o)
            if mode == "ref":
                rep["gemm_vs_ref_bit_equal_frac"] = float((y.cpu().view(torch.int16) == o.view(torch.int16)).double().mean())
    report[tag] = rep
    print(tag, json.dumps(rep), flush=True)


def glue_cases(out: dict, report: dict):
    from oracle import glue as G
    g = torch.Generator().manual_seed(77)
    for hT in (torch.bfloat16, torch.float16):
        d = str(hT).split(".")[-1]
        x = (torch.randn(2, 40, 256, generator=g) * 3).to(hT)
        w = (1 + 0.1 * torch.randn(256, generator=g)).to(hT)
        b = (0.1 * torch.randn(256, generator=g)).to(hT)
        x2 = (torch.randn(2, 40, 256, generator=g)).to(hT)
        sc = torch.randn(2, 1, 256, generator=g).to(hT)
        bi = torch.randn(2, 1, 256, generator=g).to(hT)
        xm = (torch.randn(2, 1, 6 * 64, generator=g)).to(hT)
        pre = f"glue_{d}."
        out[pre + "x"], out[pre + "w"], out[pre + "b"], out[pre + "x2"], out[pre + "sc"], out[pre + "bi"], out[pre + "xm"] = map(
            npy, (x, w, b, x2, sc, bi, xm))
        xd, wd, bd, x2d = x.to(DEV), w.to(DEV), b.to(DEV), x2.to(DEV)
        res = {
            "silu": R.glue_activation("silu", xd), "gelu": R.glue_activation("gelu", xd),
            "layernorm_affine": R.glue_layernorm(xd, wd, bd, 1e-6), "layernorm_plain": R.glue_layernorm(xd, None, None, 1e-6),
            "rms_norm": R.glue_rms_norm(xd, wd, 1e-6), "add": R.glue_add(xd, x2d),
            "mul_add_batch": R.glue_mul_add_batch(xd.clone(), sc.to(DEV), True, 1.0, bi.to(DEV), True),
            "cast_f32": R.glue_cast(xd, torch.float32),
        }
        for i, t in enumerate(R.glue_split_mod(xm.to(DEV), 6)):
            res[f"split_mod6_{i}"] = t
        torch.cuda.synchronize()
        for k, v in res.items():
            out[pre + "ref_" + k] = npy(v)
        orc = {
            "silu": G.silu(x), "gelu": G.gelu_new(x), "layernorm_affine": G.layernorm(x, w, b, 1e-6), "layernorm_plain": G.layernorm(x, None, None, 1e-6),
            "rms_norm": G.rms_norm(x, w, 1e-6), "add": G.add(x, x2), "mul_add_batch": G.mul_add_batch(x.clone(), sc, True, 1.0, bi, True),
            "cast_f32": G.cast(x, torch.float32),
        }
        rep = {}
        for k, v in orc.items():
            r = res[k].cpu()
            rep[k] = {"rel": O.rel_fro(r, v), "bit_equal_frac": float((r == v).double().mean())}
        report[f"glue_{d}"] = rep
        print(f"glue_{d}", json.dumps(rep), flush=True)


def attention_case(out: dict, report: dict):
    """test_pack_qkv -> attention_fp16: pins SURVEY row N1 (the packed layout stays private to the reference)."""
    g = torch.Generator().manual_seed(99)
    T, H = 512, 2
    qkv = (torch.randn(T, 3 * H * 128, generator=g) * 0.5).to(torch.float16)
    oq = torch.zeros(1, H, T, 128, dtype=torch.float16, device=DEV)
    ok_, ov = torch.zeros_like(oq), torch.zeros_like(oq)
    R.test_pack_qkv(qkv.to(DEV), oq, ok_, ov, T)
    for hT in (torch.float16, torch.bfloat16):
        o = torch.empty(1, T, H * 128, dtype=hT, device=DEV)
        R.attention_fp16(oq, ok_, ov, o, 128 ** -0.5)
        torch.cuda.synchronize()
        d = str(hT).split(".")[-1]
        out[f"attn.ref_o_{d}"] = npy(o)
        q, k, v = [t.double().view(T, H, 128).transpose(0, 1) for t in qkv.split(H * 128, dim=1)]
        ref = (torch.softmax(q @ k.transpose(1, 2) * 128 ** -0.5, -1) @ v).transpose(0, 1).reshape(1, T, H * 128)
        report[f"attn_{d}"] = {"vs_fp64": O.rel_fro(o.cpu(), ref)}
        print(f"attn_{d}", report[f"attn_{d}"], flush=True)
    out["attn.qkv"] = npy(qkv)
    out["attn.ref_q"], out["attn.ref_k"], out["attn.ref_v"] = npy(oq), npy(ok_), npy(ov)


def awq_case(out: dict, report: dict):
    """gemv_awq (src/kernels/awq/gemv_awq.cu:101-294): pins SURVEY row N2.  Random packed ints are valid inputs."""
    g = torch.Generator().manual_seed(55)
    n, k, gs = 256, 3072, 64
    qweight = torch.randint(-2**31, 2**31 - 1, (n // 4, k // 8 * 4), generator=g, dtype=torch.int64).to(torch.int32)
    for hT in (torch.bfloat16, torch.float16):
        d = str(hT).split(".")[-1]
        scales = (torch.rand(k // gs, n, generator=g) * 0.02 + 0.005).to(hT)
        zeros = (-(torch.rand(k // gs, n, generator=g) * 0.1 + 0.03)).to(hT)
        for m in (1, 3):
            x = torch.randn(m, k, generator=g).to(hT)
            y = R.gemv_awq(x.to(DEV), qweight.to(DEV), scales.to(DEV), zeros.to(DEV), m, n, k, gs)
            torch.cuda.synchronize()
            out[f"awq_{d}.x{m}"], out[f"awq_{d}.ref_y{m}"] = npy(x), npy(y)
        out[f"awq_{d}.scales"], out[f"awq_{d}.zeros"] = npy(scales), npy(zeros)
    out["awq.qweight"] = npy(qweight)
    report["awq"] = "stored"


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    assert R.available("ref"), "oracle/_ref/libnunchaku_ref.so missing: run oracle/ref_build/build_ref.sh in the authoring container"
    out, report = {}, {"device": torch.cuda.get_device_name(0)}
    bf, hf = torch.bfloat16, torch.float16
    run_case("plain_bf16", bf, 256, 256, 32, 300, out, report, seed=1)
    run_case("plain_fp16", hf, 256, 256, 32, 300, out, report, seed=2)
    run_case("longk_bf16", bf, 128, 3072, 32, 128, out, report, seed=3)
    run_case("longk_fp16", hf, 128, 3072, 32, 128, out, report, seed=4)
    run_case("silu_bf16", bf, 128, 256, 16, 256, out, report, seed=5, silu=True, lora_scales=[0.5])
    run_case("scales_bf16", bf, 128, 256, 32, 40, out, report, seed=6, lora_scales=[0.5, 2.0])
    run_case("glu_bf16", bf, 128, 256, 32, 200, out, report, seed=7, glu=True)
    run_case("mlp_bf16", bf, 256, 256, 32, 300, out, report, seed=8, gelu_quant_next=128)
    run_case("mlp_fp16", hf, 256, 256, 32, 300, out, report, seed=9, gelu_quant_next=128)
    run_case("rope_bf16", bf, 384, 256, 32, 256, out, report, seed=10, rope=True)
    run_case("rope_fp16", hf, 384, 256, 32, 200, out, report, seed=11, rope=True)
    glue_cases(out, report)
    attention_case(out, report)
    awq_case(out, report)
    np.savez_compressed(os.path.join(OUT_DIR, "ref_gpu_golden.npz"), **out)
    with open(os.path.join(OUT_DIR, "ref_gpu_golden_report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print("wrote", os.path.join(OUT_DIR, "ref_gpu_golden.npz"), os.path.getsize(os.path.join(OUT_DIR, "ref_gpu_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
