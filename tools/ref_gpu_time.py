"""Times the REFERENCE's own INT4 kernels (oracle/_ref/libnunchaku_ref.so, built for sm_100a) per FLUX shape on this GPU,
next to ours (INT4 and NVFP4), CUDA events, L2 flushed between iterations.

    python tools/ref_gpu_time.py [--iters 10] [--out gpurun_out/ref_gpu_time.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from op_sweep import make_layer, time_fn  # noqa: E402
from oracle import refgpu as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_gpu_time.json"))
    args = ap.parse_args()
    dev = torch.device("cuda")
    dtype = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    shapes = [(4096, 3072, 3072), (4352, 3072, 9216), (4352, 3072, 12288), (4352, 12288, 3072), (256, 3072, 3072), (256, 3072, 12288)]
    rows = []
    for (M, K, N) in shapes:
        m = make_layer(K, N, 32, "int4", dtype, dev, g)
        x = torch.randn(M, K, generator=g, device=dev).to(dtype)
        # reference: the same checkpoint tensors (random bytes in the reference layout ARE a valid checkpoint)
        act, asc, la = R.quantize_w4a4_act_fuse_lora(x, m.proj_down.data, m.smooth_factor.data)
        out = torch.empty(M, N, dtype=dtype, device=dev)
        rq, _ = time_fn(lambda: R.quantize_w4a4_act_fuse_lora(x, m.proj_down.data, m.smooth_factor.data), args.iters, flush)
        rg, _ = time_fn(lambda: R.gemm_w4a4(act, m.qweight.data, out=out, ascales=asc, wscales=m.wscales.data, lora_act_in=la,
                                            lora_up=m.proj_up.data, bias=m.bias.data), args.iters, flush)
        q, s, l2 = m.quantize(x)
        oq, _ = time_fn(lambda: m.quantize(x), args.iters, flush)
        og, _ = time_fn(lambda: m.forward_quant(q, s, l2, out), args.iters, flush)
        m4 = make_layer(K, N, 32, "nvfp4", dtype, dev, g)
        q4, s4, l4 = m4.quantize(x)
        fq, _ = time_fn(lambda: m4.quantize(x), args.iters, flush)
        fg, _ = time_fn(lambda: m4.forward_quant(q4, s4, l4, out), args.iters, flush)
        fl = 2 * M * K * N + 2 * M * 32 * N
        row = {"M": M, "K": K, "N": N, "ref_int4_quant_us": rq, "ref_int4_gemm_us": rg, "ref_int4_gemm_tflops": fl / rg / 1e6,
               "ours_int4_quant_us": oq, "ours_int4_gemm_us": og, "ours_int4_gemm_tflops": fl / og / 1e6,
               "ours_nvfp4_quant_us": fq, "ours_nvfp4_gemm_us": fg, "ours_nvfp4_gemm_tflops": fl / fg / 1e6,
               "int4_gemm_speedup_vs_ref": rg / og, "nvfp4_gemm_speedup_vs_ref_int4": rg / fg}
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
