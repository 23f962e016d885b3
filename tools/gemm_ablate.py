"""Ablation timings of the NVFP4 cluster GEMM (NB200_GEMM_DEBUG bits, results invalid, timings are the point).

    python tools/gemm_ablate.py [--M 4096 --K 3072 --N 3072] [--bn 1024,2048]
bits: 4 no main-loop MMAs | 8 no epilogue math/stores | 16 no scale-factor copies | 32 MMA warp does not wait for operands
      | 64 producer issues no loads (with 32) | 128 no low-rank branch
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

from op_sweep import make_layer, time_fn  # noqa: E402

CASES = [(0, "full kernel"), (8192, "cluster-scope acquire on the MMA warp's waits (valid)"), (1024, "3-stage TMA ring (valid)"), (2048, "2-stage TMA ring (valid)"), (512, "low-rank conversion at tile start (valid)"), (128, "no low-rank"), (16, "no SF copies"), (8, "no epilogue math/stores"), (4, "no MMAs (loads + copies + epilogue)"),
         (96, "no loads, no operand waits (MMA + copies + epilogue)"), (96 + 16, "no loads, no copies (MMA + epilogue)"),
         (96 + 16 + 128, "no loads, no copies, no low-rank"), (96 + 16 + 128 + 8, "MMA issue only"), (4 + 16 + 8 + 128, "load pipeline only")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=4096)
    ap.add_argument("--K", type=int, default=3072)
    ap.add_argument("--N", type=int, default=3072)
    ap.add_argument("--bn", default="1024,2048")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--valid-only", action="store_true", help="only the configurations whose results stay valid (A/B of schedule choices)")
    args = ap.parse_args()
    from nunchaku_b200.ops import gemm as G

    libc = ctypes.CDLL(None)
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    m = make_layer(args.K, args.N, 32, "nvfp4", torch.bfloat16, dev, g)
    x = torch.randn(args.M, args.K, generator=g, device=dev).to(torch.bfloat16)
    q, s, la = m.quantize(x)
    out = torch.empty(args.M, args.N, dtype=torch.bfloat16, device=dev)
    fl = 2 * args.M * args.K * args.N
    rows = []
    for bn in [int(b) for b in args.bn.split(",")]:
        G.BLOCK_N_OVERRIDE = bn
        for bits, name in ((CASES[:5] if args.valid_only else CASES) if bn >= 1024 else CASES[:1]):   # older kernels: control measurement on the same box
            libc.setenv(b"NB200_GEMM_DEBUG", str(bits).encode(), 1)
            print(f"bn={bn:5d} bits={bits:4d} ...", end=" ", flush=True)
            t, tmin = time_fn(lambda: m.forward_quant(q, s, la, out), args.iters, flush)
            rows.append({"bn": bn, "bits": bits, "what": name, "us": t, "min_us": tmin, "tflops_equiv": fl / t / 1e6})
            print(f"{name:55s} {t:8.1f} us  (min {tmin:.1f})", flush=True)
        libc.setenv(b"NB200_GEMM_DEBUG", b"0", 1)
    G.BLOCK_N_OVERRIDE = 0
    json.dump({"M": args.M, "K": args.K, "N": args.N, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", f"ablate_{args.M}x{args.K}x{args.N}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
