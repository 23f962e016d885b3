"""Per-shape timing of the two ops with CUDA events (L2 flushed between iterations).

    python tools/op_sweep.py [--precision int4|nvfp4|both] [--iters 20] [--out gpurun_out/sweep.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def make_layer(K, N, R, precision, dtype, dev, g):
    from nunchaku_b200.models.linear import SVDQW4A4Linear

    m = SVDQW4A4Linear(K, N, rank=R, bias=True, precision=precision, torch_dtype=dtype, device=dev)
    with torch.no_grad():
        m.qweight.copy_(torch.randint(-128, 128, m.qweight.shape, generator=g, device=dev, dtype=torch.int8))
        if precision == "nvfp4":
            m.wscales.copy_(torch.randint(0x30, 0x38, m.wscales.shape, generator=g, device=dev, dtype=torch.uint8).view(torch.float8_e4m3fn))
            m.wtscale = 1.0 / (2.6 * 0.72 * K ** 0.5)
            m.wcscales.copy_((1.0 + 0.05 * torch.randn(N, generator=g, device=dev)).to(dtype))
        else:
            m.wscales.copy_(((0.75 + 0.5 * torch.rand(m.wscales.shape, generator=g, device=dev)) / (4.6 * K ** 0.5)).to(dtype))
        m.bias.copy_((0.1 * torch.randn(N, generator=g, device=dev)).to(dtype))
        m.smooth_factor.copy_((0.75 + 0.5 * torch.rand(K, generator=g, device=dev)).to(dtype))
        m.proj_down.copy_((torch.randn(K, R, generator=g, device=dev) / K ** 0.5).to(dtype))
        m.proj_up.copy_((0.1 * torch.randn(N, R, generator=g, device=dev) / R ** 0.5).to(dtype))
    return m


def time_fn(fn, iters, flush):
    """CUDA-event time of fn() with L2 flushed before every call.  A long spin kernel is queued first
    so that the CPU enqueues the whole [flush, event, fn, event] train while the GPU is still busy:
    no launch-latency gap hides inside an event pair."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(6e7))   # ~30 ms at 1.9 GHz
    evs = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="both")
    ap.add_argument("--iters", type=int, default=15)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--shapes", default="flux")
    ap.add_argument("--bn", default="0,256,512,1024,2048")
    args = ap.parse_args()
    from nunchaku_b200.ops import gemm as G

    dev = torch.device("cuda")
    dtype = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6575.8)
    bf16 = peaks.get("bf16_tflops", 1746.5)
    R = 32
    shapes = [(4096, 3072, 3072), (4352, 3072, 9216), (4352, 3072, 12288), (4352, 12288, 3072), (256, 3072, 3072), (256, 3072, 12288), (256, 12288, 3072)]
    if args.shapes == "primary":
        shapes = shapes[:1]
    precs = ["int4", "nvfp4"] if args.precision == "both" else [args.precision]
    rows = []
    for prec in precs:
        fp4 = prec == "nvfp4"
        for (M, K, N) in shapes:
            m = make_layer(K, N, R, prec, dtype, dev, g)
            x = torch.randn(M, K, generator=g, device=dev).to(dtype)
            q, s, la = m.quantize(x)
            out = torch.empty(M, N, dtype=dtype, device=dev)
            qt, qmin = time_fn(lambda: m.quantize(x), args.iters, flush)
            Mp = q.shape[0]
            qbytes = 2 * M * K + Mp * K // 2 + (K // (16 if fp4 else 64)) * Mp * (1 if fp4 else 2) + 4 * Mp * R + 2 * K + 2 * K * R
            row = {"precision": prec, "M": M, "K": K, "N": N, "quant_us": qt, "quant_min_us": qmin, "quant_GBs": qbytes / qt / 1e3,
                   "quant_frac_hbm": qbytes / qt / 1e3 / hbm}
            for bn in [int(b) for b in args.bn.split(",")]:
                if bn and N % min(bn, 256):
                    continue
                G.BLOCK_N_OVERRIDE = bn
                try:
                    gt, gmin = time_fn(lambda: m.forward_quant(q, s, la, out), args.iters, flush)
                finally:
                    G.BLOCK_N_OVERRIDE = 0
                fl = 2 * M * K * N + 2 * M * R * N
                row[f"gemm_bn{bn}_us"] = gt
                row[f"gemm_bn{bn}_tflops"] = fl / gt / 1e6
                row[f"gemm_bn{bn}_frac"] = fl / gt / 1e6 / (bf16 * (4 if fp4 else 1))
            rows.append(row)
            print(json.dumps(row), flush=True)
            del m
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peaks": {"hbm_gbs": hbm, "bf16_tflops_burst": bf16}, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
