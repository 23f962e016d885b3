"""HBM roofline of the glue kernels (row a14): algorithmic bytes / CUDA-event time vs MEASURED_PEAKS hbm."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nunchaku_b200.ops import glue  # noqa: E402


def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()  # evict L2 (126 MB)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200000)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    peak = 6575.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbps"])
    except Exception:
        pass
    dt = torch.bfloat16
    T, C = 4352, 3072
    x = torch.randn(8, T, C, device="cuda").to(dt)   # 214 MB > L2
    y = torch.randn(8, T, C, device="cuda").to(dt)
    sc = torch.randn(8, 1, C, device="cuda").to(dt)
    sh = torch.randn(8, 1, C, device="cuda").to(dt)
    w = torch.randn(C, device="cuda").to(dt)
    x6 = torch.randn(8, T, 6 * 512, device="cuda").to(dt)
    o32 = torch.empty(8, T, C, device="cuda", dtype=torch.float32)
    nb = x.numel() * 2
    rows = []
    for name, fn, byts in [
        ("silu", lambda: glue.silu(x), 2 * nb), ("gelu_new", lambda: glue.gelu_new(x), 2 * nb),
        ("layernorm", lambda: glue.layernorm(x, None, None, 1e-6), 2 * nb), ("rms_norm", lambda: glue.rms_norm(x, w, 1e-6), 2 * nb),
        ("add", lambda: glue.add(x, y), 3 * nb), ("mul_add_batch", lambda: glue.mul_add_batch(x, sc, True, 1.0, sh, True), 2 * nb),
        ("split_mod6", lambda: glue.split_mod(x6, 6), 2 * nb), ("cast bf16->f32", lambda: glue.cast(x, o32), 3 * nb),
    ]:
        t = timeit(fn)
        rows.append({"op": name, "us": t * 1e6, "GBps": byts / t / 1e9, "frac_hbm": byts / t / 1e9 / peak})
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
