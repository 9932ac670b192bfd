"""TEST INFRASTRUCTURE (CPU oracle, never imported by the product path).

Restatement of the reference's experimental *multi-level pooled* sparse attention.  The
forward is restated here; the backward (K9:695-1237) is its plain gradient -- autograd through `multilevel_attention`
reproduces the reference's dq / dk / dv (tests/test_oracle_multilevel.py):
  N  = cogvideox/sample_evaluate/Triton/cogvideo_newattn.py
  K9 = cogvideox/sample_evaluate/Triton/kernels/block_sparse_attn_kernel_with_backward_9_10.py

A block mask entry is a LEVEL: 0 = skip the (query block, key block) pair, 1 = attend the 128 keys of the block,
L in {2, 4, 8} = attend the block's 128/L mean-pooled keys (and values) with `+log(L)` added to the scaled score, all
inside ONE softmax per query row (K9:135-277, 339-692; new_kernel.pdf Alg. 1).  Pinned against the reference's own
Triton kernel run under TRITON_INTERPRET=1 (oracle/make_golden_multilevel.py -> tests/golden/multilevel.npz).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

DEFAULT_RATIOS: Dict[int, Tuple[float, float]] = {1: (0.0, 0.05), 2: (0.05, 0.15), 4: (0.15, 0.55), 8: (0.55, 1.0)}


def multilevel_mask(attn: torch.Tensor, mask_ratios: Optional[Dict[int, Tuple[float, float]]] = None) -> torch.Tensor:
    """transfer_attn_to_mask (N:154-207): per row, rank the key blocks by score (descending; the sort is pinned to
    stable so that ties resolve like torch's CPU sort) and give rank range [int(n*a), int(n*b)) the level of that
    range; everything else is 0; the last two rows and columns are forced to level 1 (N:201-203)."""
    ratios = DEFAULT_RATIOS if mask_ratios is None else mask_ratios
    n = attn.shape[-1]
    order = torch.sort(attn, dim=-1, descending=True, stable=True).indices
    ranks_level = torch.zeros(n, dtype=torch.int32)
    for level, (a, b) in ratios.items():                       # later ranges overwrite earlier ones (N:186-199)
        lo, hi = max(0, int(n * a)), min(n, int(n * b))
        if lo < hi:
            ranks_level[lo:hi] = level
    mask = torch.zeros_like(attn, dtype=torch.int32)
    mask.scatter_(-1, order, ranks_level.expand_as(order).contiguous())
    mask[..., :, -2:] = 1
    mask[..., -2:, :] = 1
    return mask


def pad_replicate(x: torch.Tensor, multiple: int) -> torch.Tensor:
    """K9:1239-1250."""
    r = x.shape[2] % multiple
    if r:
        x = F.pad(x, (0, 0, 0, multiple - r), mode="replicate")
    return x


def pool2(x: torch.Tensor) -> torch.Tensor:
    """K9:1252-1270 with zoom_ratio 2: mean of consecutive pairs (replicate-padded to an even length), in x.dtype."""
    x = pad_replicate(x, 2)
    B, H, L, D = x.shape
    return torch.mean(x.view(B, H, L // 2, 2, D), dim=3)


def pyramid(x: torch.Tensor, block: int = 128) -> Dict[int, torch.Tensor]:
    """K9:1307-1316: level L tensor = L/2 rounds of pair pooling of the block-padded tensor (each round rounds to the
    tensor dtype, like the reference); level 1 is the UNPADDED input (the kernel masks its loads instead)."""
    p1 = pad_replicate(x, block)
    p2 = pool2(p1)
    p4 = pool2(p2)
    p8 = pool2(p4)
    return {1: x, 2: p2, 4: p4, 8: p8}


def multilevel_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, level_mask: torch.Tensor,
                         sm_scale: Optional[float] = None, block_m: int = 128, block_n: int = 128) -> torch.Tensor:
    """_fwd_kernel (K9:339-692) for BLOCK_M = BLOCK_N = POOLING_BLOCK_N = 128 (what N:10 instantiates), in fp32.

    For query block i and key block j with level L > 0 the kernel loads 128/L rows of K_L / V_L starting at
    j*128/L, with loads masked to the tensor's length -- rows beyond it read as ZERO keys and ZERO values and are NOT
    excluded from the softmax (K9:108-119,155-178): they add exp(log L - m) to the denominator.  The restatement
    keeps that behaviour ("padded key columns act as zero-score / zero-value keys")."""
    B, H, Sq, D = q.shape
    scale = (1.0 / math.sqrt(D)) if sm_scale is None else sm_scale
    kp, vp = pyramid(k, block_n), pyramid(v, block_n)
    nqb = -(-Sq // block_m)
    nkb = -(-k.shape[2] // block_n)
    assert tuple(level_mask.shape) == (B, H, nqb, nkb), (tuple(level_mask.shape), (B, H, nqb, nkb))
    out = torch.zeros(B, H, Sq, D, dtype=torch.float32)
    for b in range(B):
        for h in range(H):
            for i in range(nqb):
                rows = slice(i * block_m, min(Sq, (i + 1) * block_m))
                qi = q[b, h, rows].float()
                keys, vals, bias = [], [], []
                for j in range(nkb):
                    L = int(level_mask[b, h, i, j])
                    if L == 0:
                        continue
                    width = block_n // L
                    kl, vl = kp[L][b, h].float(), vp[L][b, h].float()
                    start = j * width
                    kt = torch.zeros(width, D)
                    vt = torch.zeros(width, D)
                    n_ok = max(0, min(width, kl.shape[0] - start))
                    if n_ok:
                        kt[:n_ok] = kl[start:start + n_ok]
                        vt[:n_ok] = vl[start:start + n_ok]
                    keys.append(kt)
                    vals.append(vt)
                    bias.append(torch.full((width,), math.log(L)))
                if not keys:
                    continue                                    # an all-zero mask row leaves the output at 0/0 upstream
                Kc, Vc, bc = torch.cat(keys), torch.cat(vals), torch.cat(bias)
                s = qi @ Kc.T * scale + bc
                out[b, h, rows] = torch.softmax(s, dim=-1) @ Vc
    return out.to(q.dtype)


def multilevel_forward(q, k, v, grid, text_length, mask_ratios, q_offsets, k_offsets, block: int = 128):
    """AdaptiveBlockSparseAttnTrain.forward of N (N:236-267): Gilbert rearrangement with the text tokens moved to
    the tail (N:129-152, same as C:141-161) -> sampled-max block scores (N:64-90 -> the Triton estimator, restated
    in asa_oracle.estimator_sampled_max) -> multi-level mask (N:154-207 with the module's ratios N:13-19) ->
    multi-level attention (K9) -> inverse rearrangement.  Returns (out, level mask, block scores)."""
    from . import asa_oracle as O
    rr = O.GilbertRearranger(grid[0], grid[1], grid[2], text_length)
    qr, kr, vr = rr.rearrange(q), rr.rearrange(k), rr.rearrange(v)
    scores = O.estimator_sampled_max(qr, kr, block, q_offsets, k_offsets)
    mask = multilevel_mask(scores, mask_ratios)
    out = multilevel_attention(qr.contiguous(), kr.contiguous(), vr.contiguous(), mask)
    return rr.reversed_rearrange(out), mask, scores
