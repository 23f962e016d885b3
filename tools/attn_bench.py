"""Attention timing at FLUX sizes (24 heads x 4608 padded tokens x 128): ours (csrc/attention.cu, tcgen05) vs the reference kernel on the same
GPU (oracle/_ref/libnunchaku_ref.so, its own packed layout) vs torch SDPA fp16 -- CUDA events, L2 flushed.

    python tools/attn_bench.py [--tokens 4352] [--heads 24]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def time_fn(fn, iters, flush):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=4352)
    ap.add_argument("--heads", type=int, default=24)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ablate", type=int, nargs="*", default=[], help="NB200_ATTN_DEBUG values to time as well (csrc/attention.cu)")
    args = ap.parse_args()
    from nunchaku_b200.ops.attention import attention_fp16
    from oracle import refgpu as R

    T, H = args.tokens, args.heads
    Tpad = (T + 255) // 256 * 256
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    qkv = (torch.randn(Tpad, 3 * H * 128, generator=g, device=dev) * 0.5).to(torch.float16)
    q = torch.zeros(1, H, Tpad, 128, dtype=torch.float16, device=dev)
    k = torch.full_like(q, float("nan"))
    v = torch.zeros_like(q)
    for t, i in ((q, 0), (k, 1), (v, 2)):
        t[0, :, :T] = qkv[:T, i * H * 128:(i + 1) * H * 128].view(T, H, 128).transpose(0, 1)
    o = torch.empty(1, Tpad, H * 128, dtype=torch.float16, device=dev)
    flops = 4.0 * H * Tpad * Tpad * 128
    res = {"tokens": T, "tokens_pad": Tpad, "heads": H, "flops": flops}
    us = time_fn(lambda: attention_fp16(q, k, v, o, 128 ** -0.5), args.iters, flush)
    res["ours_us"], res["ours_tflops"] = us, flops / us / 1e6
    print(f"ours       {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
    for bits in args.ablate:
        os.environ["NB200_ATTN_DEBUG"] = str(bits)
        usb = time_fn(lambda: attention_fp16(q, k, v, o, 128 ** -0.5), args.iters, flush)
        res[f"ablate_{bits}_us"] = usb
        print(f"  NB200_ATTN_DEBUG={bits:<3d} {usb:8.1f} us", flush=True)
    os.environ.pop("NB200_ATTN_DEBUG", None)
    attention_fp16(q, k, v, o, 128 ** -0.5)
    os.environ["NB200_ATTN_V"] = "1"   # the round's first kernel (one CTA per SM, P through shared memory), kept for this comparison
    o1 = torch.empty_like(o)
    us1 = time_fn(lambda: attention_fp16(q, k, v, o1, 128 ** -0.5), args.iters, flush)
    del os.environ["NB200_ATTN_V"]
    res["ours_v1_us"], res["v2_vs_v1_rel"] = us1, float((o[:, :T].float() - o1[:, :T].float()).norm() / o1[:, :T].float().norm())
    print(f"ours (v1)  {us1:8.1f} us  {flops / us1 / 1e6:7.1f} TFLOP/s   v2 vs v1 rel {res['v2_vs_v1_rel']:.2e}", flush=True)
    if R.available("ref"):
        rq = torch.zeros_like(q)
        rk, rv = torch.zeros_like(q), torch.zeros_like(q)
        R.test_pack_qkv(qkv, rq, rk, rv, T)
        o_ref = torch.empty_like(o)
        us = time_fn(lambda: R.attention_fp16(rq, rk, rv, o_ref, 128 ** -0.5), args.iters, flush)
        res["reference_us"], res["reference_tflops"] = us, flops / us / 1e6
        print(f"reference  {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s   (unmodified sources, sm_100a build, mma.sync)", flush=True)
        torch.cuda.synchronize()
        d = (o[:, :T].float() - o_ref[:, :T].float()).norm() / o_ref[:, :T].float().norm()
        res["ours_vs_reference_rel"] = float(d)
        print(f"ours vs reference: rel {float(d):.2e}")
    qs, ks, vs = q[:, :, :T].contiguous(), k[:, :, :T].contiguous(), v[:, :, :T].contiguous()
    fl2 = 4.0 * H * T * T * 128
    us = time_fn(lambda: torch.nn.functional.scaled_dot_product_attention(qs, ks, vs), args.iters, flush)
    res["torch_sdpa_us"], res["torch_sdpa_tflops"] = us, fl2 / us / 1e6
    print(f"torch SDPA {us:8.1f} us  {fl2 / us / 1e6:7.1f} TFLOP/s   (fp16, unpadded, library kernel)", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "attn_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
