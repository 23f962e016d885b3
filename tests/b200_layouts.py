"""Test-side pack/unpack of the B200 inter-op layouts declared in include/nunchaku_b200.h
(pure torch, any device).  Used to feed the CUDA GEMM with oracle-quantised operands and to
read back what the CUDA quantiser produced."""
from __future__ import annotations

import torch


# nibble position (0..7) of element e (0..7) inside its u32:  p = e/2 + 4*(e%2)
_NIB = [0, 4, 1, 5, 2, 6, 3, 7]


def pack_int4(q: torch.Tensor, signed: bool) -> torch.Tensor:
    """q [R, K] integer values (-8..7 signed / 0..15 unsigned) -> uint8 [R, K/2]."""
    R, K = q.shape
    v = q.to(torch.int64)
    if signed:
        v = v + 8
    assert int(v.min()) >= 0 and int(v.max()) <= 15
    v = v.view(R, K // 8, 8)
    word = torch.zeros(R, K // 8, dtype=torch.int64, device=q.device)
    for e in range(8):
        word |= v[:, :, e] << (4 * _NIB[e])
    b = torch.stack([(word >> (8 * i)) & 0xFF for i in range(4)], dim=-1).to(torch.uint8)
    return b.reshape(R, K // 2)


def unpack_int4(p: torch.Tensor, signed: bool) -> torch.Tensor:
    R, Kh = p.shape
    b = p.reshape(R, Kh // 4, 4).to(torch.int64)
    word = b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16) | (b[..., 3] << 24)
    out = torch.stack([(word >> (4 * _NIB[e])) & 0xF for e in range(8)], dim=-1).reshape(R, Kh * 2)
    if signed:
        out = out - 8
    return out.to(torch.int8)


def pack_fp4(codes: torch.Tensor) -> torch.Tensor:
    """e2m1 codes [R, K] (0..15) -> uint8 [R, K/2], low nibble = even k."""
    c = codes.to(torch.int64) & 0xF
    return (c[:, 0::2] | (c[:, 1::2] << 4)).to(torch.uint8)


def unpack_fp4(p: torch.Tensor) -> torch.Tensor:
    b = p.to(torch.int64)
    out = torch.stack([b & 0xF, (b >> 4) & 0xF], dim=-1).reshape(p.shape[0], p.shape[1] * 2)
    return out.to(torch.int8)


def _sf_tile_index(rows: int, G16: int, device) -> torch.Tensor:
    """byte offset of scale (row r, 16-group g) in tiles [rows/128][G16/4][32][16]."""
    r = torch.arange(rows, device=device).view(rows, 1)
    g = torch.arange(G16, device=device).view(1, G16)
    return ((r // 128) * (G16 // 4) + g // 4) * 512 + (r % 32) * 16 + ((r % 128) // 32) * 4 + (g % 4)


def pack_sf_tiles(scales_rg: torch.Tensor) -> torch.Tensor:
    """ue4m3 bit patterns, logical [rows, K/16] -> flat uint8 tiles (rows % 128 == 0, K % 64 == 0)."""
    rows, G16 = scales_rg.shape
    idx = _sf_tile_index(rows, G16, scales_rg.device).reshape(-1)
    out = torch.empty(rows * G16, dtype=torch.uint8, device=scales_rg.device)
    out[idx] = scales_rg.contiguous().view(torch.uint8).reshape(-1)
    return out


def unpack_sf_tiles(flat: torch.Tensor, rows: int, G16: int) -> torch.Tensor:
    idx = _sf_tile_index(rows, G16, flat.device)
    return flat.contiguous().view(torch.uint8).reshape(-1)[idx.reshape(-1)].view(rows, G16)


def lora_up_blocks(lu: torch.Tensor, cscale: torch.Tensor | None = None) -> torch.Tensor:
    """logical [N, R] -> blocks [Rp/32][N/8][4][8][8] (divided by cscale[n])."""
    N, R = lu.shape
    Rp = (R + 31) // 32 * 32
    v = torch.zeros(N, Rp, dtype=torch.float32, device=lu.device)
    v[:, :R] = lu.float()
    if cscale is not None:
        v = v / cscale.float().view(N, 1)
    v = v.to(lu.dtype).view(N // 8, 8, Rp // 32, 4, 8).permute(2, 0, 3, 1, 4).contiguous()
    return v.reshape(-1)


def lora_down_frags(ld: torch.Tensor) -> torch.Tensor:
    """logical [R, K] -> two fragment layouts back to back, each [K/32][Rp/8][lane = gq*4+t][8]:
      first  (TMA kernel, true k order): element (ks2, b, e) = Ld[8j+gq][kb*32 + 16*ks2 + 8*b + 2*t + e]
      second (register-streaming kernel): element e          = Ld[8j+gq][kb*32 + 8*t + e]"""
    R, K = ld.shape
    Rp = (R + 31) // 32 * 32
    v = torch.zeros(Rp, K, dtype=ld.dtype, device=ld.device)
    v[:R] = ld
    perm = v.view(Rp // 8, 8, K // 32, 4, 8).permute(2, 0, 1, 3, 4).contiguous()  # kb, j, gq, t, e
    true = v.view(Rp // 8, 8, K // 32, 2, 2, 4, 2).permute(2, 0, 1, 5, 3, 4, 6).contiguous()  # kb, j, gq, t, ks2, b, e
    return torch.cat([true.reshape(-1), perm.reshape(-1)])
