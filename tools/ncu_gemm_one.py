"""One launch of each GEMM variant for `ncu` (tools/ncu_summary.py reads the report).

    ncu --set full --clock-control none --import-source on -k regex:gemm -o gpurun_out/r02_gemm python tools/ncu_gemm_one.py [--precision nvfp4] [--bn 1024,512]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from op_sweep import make_layer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="nvfp4")
ap.add_argument("--bn", default="1024,512")
ap.add_argument("--shapes", default="4096x3072x3072,4352x12288x3072")
args = ap.parse_args()
from nunchaku_b200.ops import gemm as G  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
for shp in args.shapes.split(","):
    M, K, N = [int(v) for v in shp.split("x")]
    m = make_layer(K, N, 32, args.precision, torch.bfloat16, dev, g)
    x = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16)
    q, s, la = m.quantize(x)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for bn in [int(b) for b in args.bn.split(",")]:
        G.BLOCK_N_OVERRIDE = bn
        m.forward_quant(q, s, la, out)
        torch.cuda.synchronize()
