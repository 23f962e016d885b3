"""Row a4: the C++ twin of the module (include/nunchaku_b200_linear.hpp, class GEMM_W4A4 mirroring
src/Linear.h:53-120) must produce identical results (bit-identical wherever the kernels are run-to-run deterministic) to the Python mirror (SVDQW4A4Linear) on the same
checkpoint bytes: both sit on the same C ABI, so any difference is a host-side wiring bug (padding, repack
arguments, unsigned flag of the fused GELU->quantise hand-off, lora scales ...)."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from gpu_util import ref_layout_params
from oracle import svdq as O

pytestmark = pytest.mark.gpu


def _write_blob(path, tensors):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(tensors)))
        for name, t in tensors.items():
            t = t.detach().cpu().contiguous()
            raw = t.view(torch.uint8).numpy().tobytes() if t.dtype != torch.int64 else t.numpy().tobytes()
            f.write(struct.pack("<I", len(name)) + name.encode())
            f.write(struct.pack("<I", t.dim()) + struct.pack(f"<{t.dim()}q", *t.shape))
            f.write(struct.pack("<I", t.element_size()))
            f.write(raw)


def _read_blob(path, dtype):
    out = {}
    with open(path, "rb") as f:
        (count,) = struct.unpack("<I", f.read(4))
        for _ in range(count):
            (nl,) = struct.unpack("<I", f.read(4))
            name = f.read(nl).decode()
            (nd,) = struct.unpack("<I", f.read(4))
            shape = struct.unpack(f"<{nd}q", f.read(8 * nd))
            (elem,) = struct.unpack("<I", f.read(4))
            n = int(np.prod(shape)) * elem
            out[name] = torch.frombuffer(bytearray(f.read(n)), dtype=torch.uint8).view(dtype).view(*shape)
    return out


@pytest.mark.parametrize("precision", ["int4", "nvfp4"])
@pytest.mark.parametrize("hT", [torch.bfloat16, torch.float16])
def test_cpp_twin_matches_python_mirror(precision, hT, tmp_path):
    from nunchaku_b200._build import TWIN_DRIVER
    from nunchaku_b200.models.linear import SVDQW4A4Linear
    from nunchaku_b200.ops.fused import fused_gelu_mlp
    from nunchaku_b200.ops.gemm import svdq_gemm_w4a4_cuda

    assert os.path.exists(TWIN_DRIVER), "C++ twin driver not built (run __graft_entry__.build())"
    fp4 = precision == "nvfp4"
    D, H, R, M = 256, 512, 32, 300
    l1 = O.make_synthetic_layer(H, D, R, fp4=fp4, hT=hT, seed=91)
    l2 = O.make_synthetic_layer(D, H, R, fp4=fp4, hT=hT, seed=92)
    x = O.make_activations(M, D, hT, seed=93, smooth=l1.smooth)
    p1, p2 = ref_layout_params(l1), ref_layout_params(l2)

    blob = {"meta": torch.tensor([M, D, H, int(fp4), 1 if hT == torch.bfloat16 else 0], dtype=torch.int64), "x": x}
    for pre, p, layer in (("fc1.", p1, l1), ("fc2.", p2, l2)):
        blob[pre + "qweight"] = p["qweight"]
        blob[pre + "wscales"] = p["wscales"]
        blob[pre + "bias"] = p["bias"]
        blob[pre + "lora_down"] = p["proj_down"]
        blob[pre + "lora_up"] = p["proj_up"]
        blob[pre + "smooth"] = p["smooth"]
        if fp4:
            blob[pre + "wcscales"] = p["wcscales"]
            blob[pre + "wtscale"] = torch.tensor([layer.alpha], dtype=torch.float32)
    _write_blob(tmp_path / "in.blob", blob)
    pr = subprocess.run([TWIN_DRIVER, str(tmp_path / "in.blob"), str(tmp_path / "out.blob")], capture_output=True, text=True, timeout=300)
    assert pr.returncode == 0, pr.stdout + pr.stderr
    assert f"unsigned_next={int(not fp4)}" in pr.stdout
    got = _read_blob(tmp_path / "out.blob", hT)

    def mk(layer, p, K, N, unsigned):
        m = SVDQW4A4Linear(K, N, rank=R, bias=True, precision=precision, act_unsigned=unsigned, torch_dtype=hT, device="cuda")
        sd = {"qweight": p["qweight"], "wscales": p["wscales"], "bias": p["bias"], "smooth_factor": p["smooth"],
              "smooth_factor_orig": p["smooth"], "proj_down": p["proj_down"], "proj_up": p["proj_up"]}
        if fp4:
            sd["wcscales"] = p["wcscales"]
            m.wtscale = layer.alpha
        m.load_state_dict(sd)
        return m

    fc1, fc2 = mk(l1, p1, D, H, False), mk(l2, p2, H, D, not fp4)
    xd = x.cuda()
    y_plain = fc1(xd.view(1, M, D)).view(M, H)
    qx, asc, la = fc1.quantize(xd)
    y_silu = torch.empty(M, H, dtype=hT, device="cuda")
    svdq_gemm_w4a4_cuda(act=qx, wgt=fc1.qweight, out=y_silu, ascales=asc, wscales=fc1.wscales, lora_act_in=la, lora_up=fc1.proj_up,
                        bias=fc1.bias, fp4=fp4, alpha=fc1.wtscale, wcscales=fc1.wcscales, fuse_silu=True)
    y_mlp = fused_gelu_mlp(xd.view(1, M, D), fc1, fc2).view(M, D)
    torch.cuda.synchronize()
    for name, want in (("y_plain", y_plain), ("y_silu", y_silu)):
        assert torch.equal(got[name].view(torch.int16), want.cpu().view(torch.int16)), name
    # the fused fc1 epilogue accumulates fc2's low-rank hidden state with fp32 atomics across the N tiles (as the
    # reference does, lora.cuh:243-353 / SURVEY F8), so two RUNS of the same code agree only to fp32 summation order
    a, b = got["y_mlp"].double(), y_mlp.cpu().double()
    assert (a - b).norm() <= 2e-4 * b.norm(), "y_mlp"
    assert (got["y_mlp"].view(torch.int16) != y_mlp.cpu().view(torch.int16)).double().mean() <= 0.02, "y_mlp"
