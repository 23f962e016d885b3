"""TEST INFRASTRUCTURE (never imported by nunchaku_b200/): depthwise 3x3 convolution in exact math -- what dwconv_f16 computes
(/root/reference/src/kernels/dwconv.cu:202-340: NHWC input [N, H, W, C], weight [C, 3, 3, 1], stride 1, padding 1, alpha = 1, beta = 1 with
the bias as the C operand), in fp64.  The reference's CUTLASS instantiation accumulates in the 16-bit type (dwconv.cu:222-226), so it sits a few
16-bit ulps from this; parity unpinned against the reference kernel itself (it needs CUTLASS's conv headers, not built into oracle/_ref)."""
import torch


def dwconv3x3(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    """x [N, H, W, C], weight [C, 3, 3(, 1)], bias [C] or None -> fp64 [N, H, W, C]"""
    C = x.shape[-1]
    w = weight.double().reshape(C, 1, 3, 3)
    y = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w, None if bias is None else bias.double(), stride=1, padding=1, groups=C)
    return y.permute(0, 2, 3, 1).contiguous()
