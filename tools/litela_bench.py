"""SANA linear-attention QKV projection (EpilogueLiteLA): the fused GEMM epilogue against plain GEMM + nb200_litela_vk.

    python tools/litela_bench.py [--precision nvfp4] [--batch 2] [--tokens 1024] [--iters 30]

SANA-1.6B: hidden 2240 -> padded 2304 (72 heads of 32), QKV projection 2304 -> 6912.  CUDA events, L2 flushed before every call."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from op_sweep import make_layer, time_fn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="nvfp4,int4")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--tokens", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    import nunchaku_b200.ops.gemm as G
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K, heads = 2304, 72
    N = 3 * heads * 32
    M = args.batch * args.tokens
    res = {"shape": f"{args.batch} x {args.tokens} tokens, {K} -> {N} ({heads} heads)", "unit": "us (median, min)"}
    for precision in args.precision.split(","):
        fp4 = precision == "nvfp4"
        m = make_layer(K, N, 32, precision, torch.bfloat16, dev, g)
        x = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16)
        q, s, la = m.quantize(x)
        out_q = torch.empty(args.batch, args.tokens, N // 3, dtype=torch.bfloat16, device=dev)
        out_vk = torch.empty(args.batch, heads, 33, 32, dtype=torch.float32, device=dev)
        plain = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        kw = dict(act=q, wgt=m.qweight, ascales=s, wscales=m.wscales, lora_act_in=la, lora_up=m.proj_up, bias=m.bias, fp4=fp4,
                  alpha=m.wtscale if fp4 else 1.0, wcscales=m.wcscales if fp4 else None)
        outs = {}
        for route in ("fused", "split"):
            G.LITELA_FUSED = route == "fused"
            fn = lambda: svdq_gemm_w4a4_cuda(out_vk=out_vk, out_linearattn=out_q, **kw)  # noqa: E731
            med, mn = time_fn(fn, args.iters, flush)
            outs[route] = (out_q.clone(), out_vk.clone())
            res[f"{precision}_{route}"] = [round(med, 1), round(mn, 1)]
        G.LITELA_FUSED = None
        med, mn = time_fn(lambda: svdq_gemm_w4a4_cuda(out=plain, **kw), args.iters, flush)
        res[f"{precision}_plain_gemm_only"] = [round(med, 1), round(mn, 1)]
        dq = (outs["fused"][0].float() - outs["split"][0].float()).abs().max().item()
        sc = outs["split"][1].abs().amax(dim=-1, keepdim=True).clamp_min(1e-6)
        dvk = ((outs["fused"][1] - outs["split"][1]).abs() / sc).max().item()
        res[f"{precision}_fused_vs_split"] = {"max_abs_q": dq, "max_rel_vk": dvk}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
