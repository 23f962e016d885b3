"""nunchaku_b200 -- B200-native (sm_100a) SVDQuant W4A4 fused linear.

Host-side mirror of the reference operator surface for this one hot path:

    nunchaku.ops.gemm.svdq_gemm_w4a4_cuda                     -> nunchaku_b200.ops.gemm
    nunchaku.ops.quantize.svdq_quantize_w4a4_act_fuse_lora_cuda -> nunchaku_b200.ops.quantize
    nunchaku.models.linear.SVDQW4A4Linear                     -> nunchaku_b200.models.linear

The kernels live in csrc/ behind the C ABI of include/nunchaku_b200.h; importing the ops
requires the in-tree shared library (build with ``__graft_entry__.build()``).
"""
__version__ = "0.1.0"
