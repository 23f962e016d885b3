"""TEST INFRASTRUCTURE (oracle) -- closed-form restatement of the reference's packed formats.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package.  The product path (``nunchaku_b200``) never does.

The reference stores every weight-side tensor of an SVDQuant W4A4 linear pre-swizzled into
``mma.sync`` register-fragment order.  The only CPU statement of those layouts in the
reference is ``nunchaku/lora/flux/packer.py`` (torch view/permute chains); the CUDA
consumers are ``src/kernels/zgemm/gemm_base.cuh:266-348`` (load_act/load_wgt/load_wscale/
broadcast_wscale), ``gemm_w4a4.cuh:63-83`` (micro-scales) and ``lora.cuh:43-59``
(low-rank factors).  Here each layout is written as an explicit index formula
 and applied with gather/scatter -- no view/permute chains.

Pinned by ``tests/test_formats_golden.py`` against ``tests/golden/packer_formats.npz``,
which was produced by the reference's own packer (``tests/golden/make_golden.py``).

Conventions: ``g = lane // 4`` (0..7), ``t = lane % 4`` (0..3).
"""
from __future__ import annotations

import torch

__all__ = [
    "qweight_flat_index",
    "pack_qweight",
    "unpack_qweight",
    "scale_flat_index",
    "pack_group_scales",
    "unpack_group_scales",
    "pack_channel_vector",
    "unpack_channel_vector",
    "micro_scale_flat_index",
    "pack_micro_scales",
    "unpack_micro_scales",
    "lowrank_flat_index",
    "pack_lowrank",
    "unpack_lowrank",
    "pack_rotemb",
    "ref_act_flat_index",
    "pack_ref_act",
    "unpack_ref_act",
    "ref_ascale_flat_index",
    "pack_ref_ascales",
    "unpack_ref_ascales",
    "ref_lora_act_flat_index",
    "pack_ref_lora_act",
    "unpack_ref_lora_act",
]


# --------------------------------------------------------------------------------------
# A.1  qweight: int8 [N, K/2]  ==  uint32 words, 8 nibbles each
#   packer.py:187-239 (pack_weight); consumed by gemm_base.cuh:278-294 (load_wgt) and
#   gemm_w4a4.cuh:408-426 (mma: {x,y} -> rows g, {z,w} -> rows g+8 of 16-row tile j)
# --------------------------------------------------------------------------------------
def qweight_flat_index(N: int, K: int, device=None) -> tuple[torch.Tensor, torch.Tensor]:
    """For every logical (n, k) return (uint32 word index, nibble index 0..7).

    word = (((nt*(K/64) + kt)*8 + j)*32 + lane)*4 + (h*2 + c),  nibble r (bits 4r..4r+3)
    holds  W[nt*128 + j*16 + h*8 + g,  kt*64 + c*32 + t*8 + r].
    """
    assert N % 128 == 0 and K % 64 == 0, (N, K)
    n = torch.arange(N, device=device).view(N, 1)
    k = torch.arange(K, device=device).view(1, K)
    nt, n_in = n // 128, n % 128
    j, h, g = n_in // 16, (n_in % 16) // 8, n_in % 8
    kt, k_in = k // 64, k % 64
    c, t, r = k_in // 32, (k_in % 32) // 8, k_in % 8
    lane = g * 4 + t
    word = (((nt * (K // 64) + kt) * 8 + j) * 32 + lane) * 4 + (h * 2 + c)
    return word.expand(N, K), r.expand(N, K)


def pack_qweight(codes: torch.Tensor) -> torch.Tensor:
    """codes: integer tensor [N, K]; only the low 4 bits of each entry are kept (two's
    complement s4 for INT4, e2m1 code for NVFP4).  Returns int8 [N, K/2]."""
    N, K = codes.shape
    word, nib = qweight_flat_index(N, K, codes.device)
    vals = (codes.to(torch.int64) & 0xF) << (4 * nib.to(torch.int64))
    words = torch.zeros(N * K // 8, dtype=torch.int64, device=codes.device)
    words.scatter_add_(0, word.reshape(-1), vals.reshape(-1))  # disjoint nibbles: add == or
    # little-endian bytes of each 32-bit word
    b = torch.stack([(words >> (8 * i)) & 0xFF for i in range(4)], dim=-1).to(torch.uint8)
    return b.reshape(N, K // 2).view(torch.int8)


def unpack_qweight(packed: torch.Tensor, signed: bool = True) -> torch.Tensor:
    """packed: int8/uint8 [N, K/2] in reference layout -> int8 [N, K] (sign-extended s4 if
    ``signed`` else raw 0..15 codes)."""
    N, Kh = packed.shape
    K = Kh * 2
    b = packed.contiguous().view(torch.uint8).reshape(-1, 4).to(torch.int64)
    words = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16) | (b[:, 3] << 24)
    word, nib = qweight_flat_index(N, K, packed.device)
    v = (words[word.reshape(-1)] >> (4 * nib.reshape(-1).to(torch.int64))) & 0xF
    if signed:
        v = torch.where(v >= 8, v - 16, v)
    return v.reshape(N, K).to(torch.int8)


# --------------------------------------------------------------------------------------
# A.2  INT4 group scales, nominal [K/64, N]; packer.py:241-301 (pack_scale), consumed by
#   gemm_base.cuh:317-348 (load_wscale / broadcast_wscale)
# --------------------------------------------------------------------------------------
def scale_flat_index(N: int, G: int, device=None) -> torch.Tensor:
    """flat index for logical (n, grp):  [nt][grp][lane = a*4 + c2][b*2 + d]  with
    n = nt*128 + a*16 + b*8 + c2*2 + d."""
    assert N % 128 == 0
    n = torch.arange(N, device=device).view(N, 1)
    grp = torch.arange(G, device=device).view(1, G)
    nt, n_in = n // 128, n % 128
    a, b, c2, d = n_in // 16, (n_in % 16) // 8, (n_in % 8) // 2, n_in % 2
    lane = a * 4 + c2
    return ((nt * G + grp) * 32 + lane) * 4 + (b * 2 + d)


def pack_group_scales(scales_ng: torch.Tensor) -> torch.Tensor:
    """scales_ng: logical [N, G] (fp16/bf16) -> packed tensor of nominal shape [G, N]."""
    N, G = scales_ng.shape
    idx = scale_flat_index(N, G, scales_ng.device).reshape(-1)
    out = torch.empty(N * G, dtype=scales_ng.dtype, device=scales_ng.device)
    out[idx] = scales_ng.reshape(-1)
    return out.view(G, N)


def unpack_group_scales(packed: torch.Tensor, N: int | None = None) -> torch.Tensor:
    """packed nominal [G, N] -> logical [N, G]."""
    G, N_ = packed.shape
    N = N or N_
    idx = scale_flat_index(N, G, packed.device)
    return packed.reshape(-1)[idx.reshape(-1)].view(N, G)


def pack_channel_vector(v: torch.Tensor) -> torch.Tensor:
    """bias / smooth_factor / wcscales: ``pack_scale(group_size=-1)`` i.e. one group
    (nunchaku_converter.py:944-945; EpilogueBias gemm_base.cuh:713-731; smooth in
    EpilogueQuantize gemm_w4a4.cuh:958-959)."""
    return pack_group_scales(v.reshape(-1, 1)).reshape(-1)


def unpack_channel_vector(packed: torch.Tensor) -> torch.Tensor:
    return unpack_group_scales(packed.reshape(1, -1)).reshape(-1)


# --------------------------------------------------------------------------------------
# A.2  NVFP4 micro-scales (fp8 e4m3, one per 16 k), nominal [K/16, N];
#   packer.py:303-360 (pack_micro_scale), consumed by gemm_w4a4.cuh:63-83 (load_wmscale)
# --------------------------------------------------------------------------------------
def micro_scale_flat_index(N: int, G16: int, device=None) -> torch.Tensor:
    """flat index for logical (n, g16): [nt][kt][lane = s*4 + q][p][kk] with
    n = nt*128 + p*32 + q*8 + s and g16 = kt*4 + kk."""
    assert N % 128 == 0 and G16 % 4 == 0
    n = torch.arange(N, device=device).view(N, 1)
    g16 = torch.arange(G16, device=device).view(1, G16)
    nt, n_in = n // 128, n % 128
    p, q, s = n_in // 32, (n_in % 32) // 8, n_in % 8
    kt, kk = g16 // 4, g16 % 4
    lane = s * 4 + q
    return ((((nt * (G16 // 4) + kt) * 32 + lane) * 4 + p) * 4) + kk


def pack_micro_scales(scales_ng: torch.Tensor) -> torch.Tensor:
    """scales_ng: logical [N, K/16] float8_e4m3fn (or uint8 bit patterns) -> [K/16, N]."""
    N, G16 = scales_ng.shape
    idx = micro_scale_flat_index(N, G16, scales_ng.device).reshape(-1)
    src = scales_ng.contiguous().view(torch.uint8).reshape(-1)
    out = torch.empty(N * G16, dtype=torch.uint8, device=scales_ng.device)
    out[idx] = src
    return out.view(G16, N).view(scales_ng.dtype if scales_ng.dtype != torch.uint8 else torch.uint8)


def unpack_micro_scales(packed: torch.Tensor) -> torch.Tensor:
    """packed nominal [K/16, N] (float8_e4m3fn or uint8) -> logical [N, K/16] same dtype."""
    G16, N = packed.shape
    idx = micro_scale_flat_index(N, G16, packed.device)
    flat = packed.contiguous().view(torch.uint8).reshape(-1)
    out = flat[idx.reshape(-1)].view(N, G16)
    return out.view(packed.dtype) if packed.dtype != torch.uint8 else out


# --------------------------------------------------------------------------------------
# A.3  low-rank factors; packer.py:362-437 and nunchaku_converter.py:71-141,
#   consumed by lora.cuh:43-59 (load_lora_wgt)
#   proj_up  : stored [N, R];  logical Lu[n, r]
#   proj_down: stored [K, R];  logical Ld[r, k]   (the *transpose* is what is tiled)
# --------------------------------------------------------------------------------------
def lowrank_flat_index(C: int, R: int, down: bool, device=None) -> torch.Tensor:
    """Flat storage index of logical element.

    up   (down=False): logical [N=C, R]  -> idx[n, r]
         [N/16][R/16][lane][h][c][e]  with n = 16*i + h*8 + g,  r = 16*u + c*8 + t*2 + e
    down (down=True) : logical [R, K=C]  -> idx[r, k]
         [K/16][R/16][lane][h][c][e]  with r = 16*u + h*8 + g,  k = 16*i + c*8 + t*2 + e
    """
    assert C % 16 == 0 and R % 16 == 0, (C, R)
    if not down:
        n = torch.arange(C, device=device).view(C, 1)
        r = torch.arange(R, device=device).view(1, R)
        i, h, g = n // 16, (n % 16) // 8, n % 8
        u, c, t, e = r // 16, (r % 16) // 8, (r % 8) // 2, r % 2
    else:
        r = torch.arange(R, device=device).view(R, 1)
        k = torch.arange(C, device=device).view(1, C)
        u, h, g = r // 16, (r % 16) // 8, r % 8
        i, c, t, e = k // 16, (k % 16) // 8, (k % 8) // 2, k % 2
    lane = g * 4 + t
    return ((((i * (R // 16) + u) * 32 + lane) * 2 + h) * 2 + c) * 2 + e


def pack_lowrank(logical: torch.Tensor, down: bool) -> torch.Tensor:
    """up: logical [N, R] -> stored [N, R];  down: logical [R, K] -> stored [K, R]."""
    if down:
        R, C = logical.shape
    else:
        C, R = logical.shape
    idx = lowrank_flat_index(C, R, down, logical.device).reshape(-1)
    out = torch.empty(C * R, dtype=logical.dtype, device=logical.device)
    out[idx] = logical.reshape(-1)
    return out.view(C, R)


def unpack_lowrank(stored: torch.Tensor, down: bool) -> torch.Tensor:
    """stored [C, R] -> logical ([N, R] for up, [R, K] for down)."""
    C, R = stored.shape
    idx = lowrank_flat_index(C, R, down, stored.device)
    return stored.reshape(-1)[idx.reshape(-1)].view(idx.shape)


# --------------------------------------------------------------------------------------
# A.4  packed rotary table; nunchaku/models/embeddings.py (pack_rotemb) and
#   nunchaku/models/transformers/transformer_flux.py:60-92; consumed by EpilogueRMSNormRope::load_rotemb
#   (epilogues.cuh:283-300): fp32 accumulator-fragment order, one float4 = {sin,cos of row g, sin,cos of row g+8}
# --------------------------------------------------------------------------------------
def pack_rotemb(sin: torch.Tensor, cos: torch.Tensor) -> torch.Tensor:
    """sin, cos: logical [M, 64] fp32 (pair i rotates head columns 2i, 2i+1) -> packed [M, 128] fp32.

    float index of (m, pair p, s) with s = 0 (sin) / 1 (cos):
        (((((m/16)*16 + p/4)*8 + m%8)*4 + p%4)*2 + (m%16)/8)*2 + s
    """
    M, P = sin.shape
    assert P == 64 and M % 16 == 0
    m = torch.arange(M, device=sin.device).view(M, 1, 1)
    p = torch.arange(64, device=sin.device).view(1, 64, 1)
    s = torch.arange(2, device=sin.device).view(1, 1, 2)
    idx = (((((m // 16) * 16 + p // 4) * 8 + m % 8) * 4 + p % 4) * 2 + (m % 16) // 8) * 2 + s
    out = torch.empty(M * 128, dtype=torch.float32, device=sin.device)
    out[idx.reshape(-1)] = torch.stack([sin, cos], dim=-1).to(torch.float32).reshape(-1)
    return out.view(M, 128)


# --------------------------------------------------------------------------------------
# A.4  The reference's INTER-OP tensors (INT4 path): what its quantizer / fused epilogue write and
#   its GEMM reads.  Ours are laid out differently (include/nunchaku_b200.h); these maps exist so
#   that outputs of the reference's kernels RUN ON THE B200 (oracle/_ref, tests/golden/ref_gpu_*.npz)
#   can be compared element by element with the oracle.  Pinned by tests/test_ref_gpu_golden.py:
#   a wrong formula would scramble ~100 % of the codes, the goldens agree to <= 2e-3.
#
#   act     uint4 [Mp/256][K/64][warp 8][mt 2][lane 32], components {x,y,z,w} =
#           (row g, k t*8+r) (row g+8, k t*8+r) (row g, k 32+t*8+r) (row g+8, k 32+t*8+r), nibble r
#           (quantize_w4a4_from_fpsum_warp gemm_w4a4.cuh:429-523, store :1002, load_act gemm_base.cuh:266-276)
#   ascales hT    [Mp/256][K/64][warp 8][lane 16][2]: row (lane/8)*16 + lane%8 + 8*e of the warp's 32
#           (pack_ascales gemm_base.cuh:452-468, load_ascale :296-312)
#   lora_act f32  [Mp/256][R/16][warp 8][mt 2][j 8][lane 32]: m16n8k16 C fragments, j = 4*nhalf + 2*h + e ->
#           row g + 8*h, rank column 8*nhalf + 2*t + e     (lora.cuh:61-94, mma gemm_base.cuh:203-227)
# --------------------------------------------------------------------------------------
def ref_act_flat_index(Mp: int, K: int, device=None) -> tuple[torch.Tensor, torch.Tensor]:
    """(uint32 word index, nibble) of logical (m, k) inside the reference's packed activation tensor."""
    assert Mp % 256 == 0 and K % 64 == 0, (Mp, K)
    m = torch.arange(Mp, device=device).view(Mp, 1)
    k = torch.arange(K, device=device).view(1, K)
    bm, m_in = m // 256, m % 256
    warp, mt, h, g = m_in // 32, (m_in % 32) // 16, (m_in % 16) // 8, m_in % 8
    kt, k_in = k // 64, k % 64
    c, t, r = k_in // 32, (k_in % 32) // 8, k_in % 8
    lane = g * 4 + t
    word = ((((bm * (K // 64) + kt) * 8 + warp) * 2 + mt) * 32 + lane) * 4 + (c * 2 + h)
    return word.expand(Mp, K), r.expand(Mp, K)


def pack_ref_act(codes: torch.Tensor) -> torch.Tensor:
    """codes: integer [Mp, K] (low 4 bits kept) -> uint8 [Mp, K/2] in the reference's activation order."""
    Mp, K = codes.shape
    word, nib = ref_act_flat_index(Mp, K, codes.device)
    vals = (codes.to(torch.int64) & 0xF) << (4 * nib.to(torch.int64))
    words = torch.zeros(Mp * K // 8, dtype=torch.int64, device=codes.device)
    words.scatter_add_(0, word.reshape(-1), vals.reshape(-1))
    b = torch.stack([(words >> (8 * i)) & 0xFF for i in range(4)], dim=-1).to(torch.uint8)
    return b.reshape(Mp, K // 2)


def unpack_ref_act(packed: torch.Tensor, signed: bool = True) -> torch.Tensor:
    """uint8/int8 [Mp, K/2] written by the reference's quantizer -> int8 [Mp, K] codes."""
    Mp, Kh = packed.shape
    K = Kh * 2
    b = packed.contiguous().view(torch.uint8).reshape(-1, 4).to(torch.int64)
    words = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16) | (b[:, 3] << 24)
    word, nib = ref_act_flat_index(Mp, K, packed.device)
    v = (words[word.reshape(-1)] >> (4 * nib.reshape(-1).to(torch.int64))) & 0xF
    if signed:
        v = torch.where(v >= 8, v - 16, v)
    return v.reshape(Mp, K).to(torch.int8)


def ref_ascale_flat_index(Mp: int, G: int, device=None) -> torch.Tensor:
    """flat element index of logical (m, group) in the reference's packed activation scales."""
    assert Mp % 256 == 0
    m = torch.arange(Mp, device=device).view(Mp, 1)
    grp = torch.arange(G, device=device).view(1, G)
    bm, m_in = m // 256, m % 256
    warp, r32 = m_in // 32, m_in % 32
    lane = (r32 // 16) * 8 + r32 % 8
    e = (r32 % 16) // 8
    return (((bm * G + grp) * 8 + warp) * 16 + lane) * 2 + e


def pack_ref_ascales(scales_mg: torch.Tensor) -> torch.Tensor:
    """logical [Mp, G] -> nominal [G, Mp] in the reference's order."""
    Mp, G = scales_mg.shape
    idx = ref_ascale_flat_index(Mp, G, scales_mg.device).reshape(-1)
    out = torch.empty(Mp * G, dtype=scales_mg.dtype, device=scales_mg.device)
    out[idx] = scales_mg.reshape(-1)
    return out.view(G, Mp)


def unpack_ref_ascales(packed: torch.Tensor) -> torch.Tensor:
    """nominal [G, Mp] written by the reference -> logical [Mp, G]."""
    G, Mp = packed.shape
    idx = ref_ascale_flat_index(Mp, G, packed.device)
    return packed.reshape(-1)[idx.reshape(-1)].view(Mp, G)


def ref_lora_act_flat_index(Mp: int, R: int, device=None) -> torch.Tensor:
    """flat f32 index of logical (m, r) in the reference's low-rank hidden state."""
    assert Mp % 256 == 0 and R % 16 == 0
    m = torch.arange(Mp, device=device).view(Mp, 1)
    r = torch.arange(R, device=device).view(1, R)
    bm, m_in = m // 256, m % 256
    warp, mt, h, g = m_in // 32, (m_in % 32) // 16, (m_in % 16) // 8, m_in % 8
    rt, r_in = r // 16, r % 16
    nhalf, t, e = r_in // 8, (r_in % 8) // 2, r_in % 2
    j = nhalf * 4 + h * 2 + e
    lane = g * 4 + t
    return ((((bm * (R // 16) + rt) * 8 + warp) * 2 + mt) * 8 + j) * 32 + lane


def pack_ref_lora_act(la: torch.Tensor) -> torch.Tensor:
    Mp, R = la.shape
    idx = ref_lora_act_flat_index(Mp, R, la.device).reshape(-1)
    out = torch.empty(Mp * R, dtype=la.dtype, device=la.device)
    out[idx] = la.reshape(-1)
    return out.view(Mp, R)


def unpack_ref_lora_act(packed: torch.Tensor) -> torch.Tensor:
    Mp, R = packed.shape
    idx = ref_lora_act_flat_index(Mp, R, packed.device)
    return packed.reshape(-1)[idx.reshape(-1)].view(Mp, R)
