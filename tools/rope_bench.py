"""The QKV projection with RMSNorm + RoPE at FLUX's single-block shape (4352 x 3072 -> 9216, NVFP4): quantize + plain cluster GEMM + csrc/rope.cu,
timed with both lane mappings of the rope kernel (NB200_ROPE=1: 8 lanes per head, 2 x 16 B per lane; 2: 16 lanes per head, 1 x 16 B per lane) and
without the rope kernel (plain projection).  CUDA events, L2 flushed before every call (the rope kernel itself reads what the GEMM just wrote, as in
the model).  Also usable under ncu:  ncu --set full -k regex:rope -s 2 -c 1 python tools/rope_bench.py --iters 1

    python tools/rope_bench.py [--iters 30]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from op_sweep import make_layer, time_fn  # noqa: E402

from nunchaku_b200.utils import pack_rotemb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
M, K, N = 4352, 3072, 9216
m = make_layer(K, N, 32, "nvfp4", torch.bfloat16, dev, g)
x = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16)
nq = (1.0 + 0.1 * torch.randn(128, generator=g, device=dev)).to(torch.bfloat16)
ang = torch.rand(M, 64, generator=g, device=dev) * 6.28
rot = pack_rotemb(torch.sin(ang), torch.cos(ang))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
q, s, la = m.quantize(x)
res = {"shape": [M, K, N], "unit": "us (median, min): GEMM + rope kernel, activations already quantised"}
outs = {}
for variant in ("1", "2"):
    os.environ["NB200_ROPE"] = variant
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    from nunchaku_b200.ops.gemm import gemm_b200

    fn = lambda: gemm_b200(q, s, la, m.b200(), out=out, norm_q=nq, norm_k=nq, rotary_emb=rot)  # noqa: E731
    med, mn = time_fn(fn, args.iters, flush)
    res[f"rope_variant_{variant}"] = [round(med, 1), round(mn, 1)]
    outs[variant] = out
os.environ.pop("NB200_ROPE")
plain = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
med, mn = time_fn(lambda: m.forward_quant(q, s, la, plain), args.iters, flush)
res["plain_gemm"] = [round(med, 1), round(mn, 1)]
res["variants_bit_equal"] = bool(torch.equal(outs["1"].view(torch.int16), outs["2"].view(torch.int16)))
print(json.dumps(res))
