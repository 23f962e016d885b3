"""Where the activation quantizer's time goes: NB200_QUANT_DEBUG ablation bits (csrc/quantize_v2.cu), cold (L2 flushed) and warm
(input left in L2 by a preceding write, as in the model) -- CUDA events around single launches.

    python tools/quant_ablate.py [--precision nvfp4|int4] [--iters 10]
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from op_sweep import make_layer  # noqa: E402

CASES = [(0, "full kernel"), (1, "no low-rank MMAs"), (4, "no code / scale stores"), (2, "no quantise phases (loads + low-rank only)"), (3, "loads only"),
         (8, "no activation loads (compute + stores)"), (8 + 4, "no loads, no stores (compute only)"), (8 + 4 + 1, "quantise math only"),
         (8 + 3, "launch + fragment ring + reduction only"), (16, "launch + barrier setup only"), (32, "launch only (kernel returns at once)")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="nvfp4")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--shapes", default="4096x3072,4352x12288,256x3072")
    args = ap.parse_args()
    libc = ctypes.CDLL(None)
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    for shp in args.shapes.split(","):
        M, K = (int(v) for v in shp.split("x"))
        m = make_layer(K, 3072, 32, args.precision, torch.bfloat16, dev, g)
        x = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16)
        m.quantize(x)
        for bits, name in CASES:
            libc.setenv(b"NB200_QUANT_DEBUG", str(bits).encode(), 1)
            res = {}
            for mode in ("cold", "warm"):
                ts = []
                for _ in range(args.iters + 2):
                    if mode == "cold":
                        flush.zero_()
                    else:
                        x.add_(0)     # rewrites x: L2 resident, like the LayerNorm / GEMM epilogue that produced it
                    torch.cuda._sleep(200000)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    m.quantize(x)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts = sorted(ts[2:])
                res[mode] = ts[len(ts) // 2]
            rows.append({"M": M, "K": K, "bits": bits, "what": name, **res})
            print(f"M={M:5d} K={K:5d} bits={bits:3d} {name:45s} cold {res['cold']:7.1f} us   warm {res['warm']:7.1f} us", flush=True)
        libc.setenv(b"NB200_QUANT_DEBUG", b"0", 1)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"quant_ablate_{args.precision}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
