"""Synthetic inputs and the algorithmic-FLOP accounting of BASELINE.md section 3.
Shared by bench.py, the tools and (through a re-export) the test oracle; no kernels, no oracle code."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def synth_qkv(B, H, S, D, seed, dtype=torch.bfloat16, structured: float = 0.0,
              grid: Optional[Tuple[int, int, int]] = None, text_length: int = 0, ramp: bool = False):
    """Gaussian q,k,v, optionally with the structured positional component of SURVEY 8(d):
    per head F ~ N(0,6^2) [3, D/2]; phase = x/W*F0 + y/H*F1 + z/T*F2 over raster coordinates;
    base = [cos phase ; sin phase]; q = a*base + N(0,1), k = a*base + N(0,1), v ~ N(0,1).
    `ramp`: the amplitude grows linearly with the frame index (a * z/T): early frames look Gaussian (the selection
    saturates at max_retain), late frames are strongly structured (min_retain), the frames between cover the range --
    the "mixed" bench input, whose per-row retained counts differ inside every head."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, H, S, D, generator=g)
    k = torch.randn(B, H, S, D, generator=g)
    v = torch.randn(B, H, S, D, generator=g)
    if structured and grid is not None:
        Wd, Ht, Dp = grid
        n_vid = Wd * Ht * Dp
        r = torch.arange(n_vid)
        x = (r % Wd).float() / Wd
        y = ((r // Wd) % Ht).float() / Ht
        z = (r // (Wd * Ht)).float() / Dp
        Fq = torch.randn(H, 3, D // 2, generator=g) * 6.0
        phase = x[None, :, None] * Fq[:, 0:1] + y[None, :, None] * Fq[:, 1:2] + z[None, :, None] * Fq[:, 2:3]
        base = torch.cat([torch.cos(phase), torch.sin(phase)], dim=-1)       # [H, n_vid, D]
        amp = structured * (z[None, :, None] if ramp else 1.0)
        q[:, :, text_length:text_length + n_vid] += (amp * base)[None]
        k[:, :, text_length:text_length + n_vid] += (amp * base)[None]
    return q.to(dtype), k.to(dtype), v.to(dtype)


def attention_flops(mask_counts_cols: torch.Tensor, S: int, D: int, block: int, n_pooled: int) -> float:
    """4*D*sum_i rows_i*(sum_{j in sel(i)} cols_j + n_pooled), ragged tail
    tiles at true size.  `mask_counts_cols` [B,H,nq] = selected key columns per q-block row."""
    nq = mask_counts_cols.size(-1)
    rows = torch.full((nq,), block, dtype=torch.float64)
    if S % block:
        rows[-1] = S % block
    tot = (rows.view(1, 1, nq) * (mask_counts_cols.double() + n_pooled)).sum()
    return float(4.0 * D * tot)
