"""Barrier-wait cycle breakdown of the fused GEMM (device-side clock64 counters, p.prof).

    python tools/gemm_prof.py [--precision int4|nvfp4|both] [--M 4096 --K 3072 --N 3072] [--bn 0,512]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from op_sweep import make_layer  # noqa: E402

SLOTS = {9: "KERNEL total", 12: "setup (alloc, barriers, sync)", 13: "mma: first operands ready", 0: "producer wait empty", 1: "mma wait tmem_empty", 2: "mma wait operands(full/cfull)", 3: "mma wait lora",
         10: "mma loop total", 4: "epi wait tmem_full", 11: "epi pre-tile work (bias+lora cvt)", 5: "epi loop total",
         6: "conv wait TMA(full)", 7: "conv wait cempty", 8: "conv busy"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="both")
    ap.add_argument("--M", type=int, default=4096)
    ap.add_argument("--K", type=int, default=3072)
    ap.add_argument("--N", type=int, default=3072)
    ap.add_argument("--bn", default="256,512,1024,2048")
    ap.add_argument("--timeline", action="store_true", help="print block 0's MMA-thread timeline (cluster kernel)")
    ap.add_argument("--fused", action="store_true", help="profile the fc1 -> GELU -> quantise-for-fc2 epilogue (EPI_QUANT)")
    args = ap.parse_args()
    from nunchaku_b200.ops import gemm as G

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    for prec in (["int4", "nvfp4"] if args.precision == "both" else [args.precision]):
        m = make_layer(args.K, args.N, 32, prec, torch.bfloat16, dev, g)
        x = torch.randn(args.M, args.K, generator=g, device=dev).to(torch.bfloat16)
        q, s, la = m.quantize(x)
        out = torch.empty(args.M, args.N, dtype=torch.bfloat16, device=dev)
        run = lambda: m.forward_quant(q, s, la, out)  # noqa: E731
        if args.fused:
            fp4 = prec == "nvfp4"
            m2 = make_layer(args.N, args.K, 32, prec, torch.bfloat16, dev, g)
            Mp = q.shape[0]
            q2 = torch.empty(Mp, args.N // 2, dtype=torch.uint8, device=dev)
            s2 = (torch.empty(args.N // 16, Mp, dtype=torch.float8_e4m3fn, device=dev) if fp4
                  else torch.empty(args.N // 64, Mp, dtype=torch.bfloat16, device=dev))
            la2 = torch.empty(Mp, 32, dtype=torch.float32, device=dev)
            run = lambda: G.svdq_gemm_w4a4_cuda(  # noqa: E731
                act=q, wgt=m.qweight, qout=q2, ascales=s, wscales=m.wscales, oscales=s2, lora_act_in=la, lora_up=m.proj_up,
                lora_down=m2.proj_down, lora_act_out=la2, bias=m.bias, smooth_factor=m2.smooth_factor, fp4=fp4, alpha=m.wtscale,
                wcscales=m.wcscales)
        for bn in [int(b) for b in args.bn.split(",")]:
            G.BLOCK_N_OVERRIDE = bn
            for _ in range(3):
                run()
            prof = torch.zeros(148 + 12 + 16 + 16, 16, dtype=torch.int64, device=dev)
            G.PROF_BUFFER = prof
            torch.cuda.synchronize()
            torch.cuda._sleep(int(2e7))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            G.PROF_BUFFER = None
            G.BLOCK_N_OVERRIDE = 0
            us = e0.elapsed_time(e1) * 1e3
            if args.timeline:
                tl = prof.cpu()[160:176].reshape(-1)
                tl = tl[tl != 0]
                names = {1: "tile begin", 2: "tmem_empty ok", 3: "operands ok", 4: "stage issued"}
                prev = 0
                out = []
                for v in tl.tolist():
                    t, tag = v >> 4, v & 15
                    out.append(f"{names.get(tag, tag)}@{t}(+{t - prev})")
                    prev = t
                print("   timeline:", "  ".join(out))
                etl = prof.cpu()[176:192].reshape(-1)
                etl = etl[etl != 0]
                enames = {1: "pre-tile done", 2: "tmem_full ok", 3: "accumulator in registers, tmem_empty sent", 4: "next tile's low-rank converted", 5: "staging free",
                          6: "chunk computed", 7: "chunk handed to TMA"}
                prev = 0
                out = []
                for v in etl.tolist():
                    t, tag = v >> 4, v & 15
                    out.append(f"{enames.get(tag, tag)}@{t}(+{t - prev})")
                    prev = t
                print("   epilogue timeline:", "  ".join(out))
            pr = prof.cpu()[:148].double()
            used = pr[:, 10] > 0 if bn < 512 else pr[:, 5] > 0
            print(f"== {prec} M={args.M} K={args.K} N={args.N} bn={bn}: {us:.1f} us  ({2*args.M*args.K*args.N/us/1e6:.0f} TFLOP/s)  CTAs with data {int(used.sum())}")
            ns = pr[:, 14]
            if (ns > 0).any():
                sel = ns > 0
                print(f"   effective SM clock inside the kernel: {(pr[sel, 9] / ns[sel]).mean().item() * 1e3:.0f} MHz   CTA wall time mean {ns[sel].mean().item() / 1e3:.1f} us max {ns[sel].max().item() / 1e3:.1f} us")
            for k, name in SLOTS.items():
                col = pr[:, k]
                nz = col[col > 0]
                if nz.numel():
                    print(f"   {name:36s} mean {nz.mean().item()/1e3:9.1f} kclk   max {nz.max().item()/1e3:9.1f} kclk   ({nz.numel()} CTAs)")


if __name__ == "__main__":
    main()
