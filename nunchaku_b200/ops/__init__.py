"""Operator layer mirroring nunchaku/ops/{gemm,quantize,fused}.py for the SVDQuant path, plus the
elementwise glue of src/kernels/misc_kernels.h (ops.glue)."""
from .gemm import svdq_gemm_w4a4_cuda  # noqa: F401
from .quantize import svdq_quantize_w4a4_act_fuse_lora_cuda  # noqa: F401
from . import glue  # noqa: F401,E402
