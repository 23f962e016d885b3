"""Operator layer mirroring nunchaku/ops/{gemm,quantize,fused}.py for the SVDQuant path."""
from .gemm import svdq_gemm_w4a4_cuda  # noqa: F401
from .quantize import svdq_quantize_w4a4_act_fuse_lora_cuda  # noqa: F401
