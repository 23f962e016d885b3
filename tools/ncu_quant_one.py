"""One launch of the activation quantizer per shape for `ncu`.

    ncu --set full --clock-control none --import-source on -k regex:quantize_v2 -o gpurun_out/r02_quant python tools/ncu_quant_one.py [--precision nvfp4]
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
ap.add_argument("--shapes", default="4096x3072,4352x12288")
args = ap.parse_args()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
for shp in args.shapes.split(","):
    M, K = [int(v) for v in shp.split("x")]
    m = make_layer(K, 3072, 32, args.precision, torch.bfloat16, dev, g)
    x = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16)
    m.quantize(x)
    torch.cuda.synchronize()
