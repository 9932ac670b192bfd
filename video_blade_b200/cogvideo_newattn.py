"""Drop-in mirror of the reference's experimental multi-level module
cogvideox/sample_evaluate/Triton/cogvideo_newattn.py (N) and of the public entry of its Triton kernels
(kernels/block_sparse_attn_kernel_with_backward_9_10.py, K9) -- SURVEY.md 8(f) rank 4.  Same module-level knobs
(`mask_ratios`, `use_rearrange`, `width/height/depth`, `text_length`, N:10-25), same names and signatures; B200 kernels
underneath (include/blade_asa.h: blade_multilevel_*), forward only.

    reference (N / K9)                                       here
    transfer_attn_to_mask(attn, mask_ratios)   (N:154-207)     blade_multilevel_mask (rank table built on the host)
    sparse_attention_fn(q, k, v, mask, None)   (N:9, K9:1578)  blade_level_mask_to_index + pyramid + attention
    adaptive_block_sparse_attn(q, k, v)        (N:210-235)     gather -> sampled-max scores -> mask -> pyramid -> attention
    AdaptiveBlockSparseAttnTrain.forward       (N:237-267)     the same, with the Gilbert gather / inverse permutation
                                                               fused into the gather kernel and the attention epilogue
"""
from __future__ import annotations

import ctypes as C
import math
import sys
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, current_stream, ptr, tensor_desc
from .asa import AsaEngine, AsaKnobs, _on_tensor_device, require_no_grad

# ----------------------------- parameters (N:10-25) -----------------------------
mask_ratios: Dict[int, Tuple[float, float]] = {
    1: (0.0, 0.05),
    2: (0.05, 0.15),
    4: (0.15, 0.25),
    8: (0.25, 0.5),
    0: (0.5, 1.0),
}
use_rearrange = True
width = 45
height = 30
depth = 13
text_length = 226
# literals of the reference: block 128 (N:9), num_keep 32 (N:64), forced last two rows / columns (N:201-203)
block_size = 128
num_keep = 32
force_last = 2

DEFAULT_RATIOS = {1: (0.0, 0.05), 2: (0.05, 0.15), 4: (0.15, 0.55), 8: (0.55, 1.0)}      # N:172-178


def rank_levels(n: int, ratios: Optional[Dict[int, Tuple[float, float]]] = None) -> np.ndarray:
    """Level of the key block ranked p-th in a row of n blocks (N:186-199): ranges are applied in dict order, later
    ranges overwrite earlier ones, everything uncovered is 0."""
    ratios = DEFAULT_RATIOS if ratios is None else ratios
    lv = np.zeros(n, np.uint8)
    for level, (a, b) in ratios.items():
        if int(level) not in (0, 1, 2, 4, 8):
            raise ValueError(f"mask level {level} not in {{0, 1, 2, 4, 8}}")
        lo, hi = max(0, int(n * a)), min(n, int(n * b))
        if lo < hi:
            lv[lo:hi] = level
    return lv


class MultiLevelEngine:
    """Device tables + workspace for the multi-level path; every tensor op of N / K9 is a kernel behind the C ABI."""

    def __init__(self):
        self.lib = _lib.load()
        self._rank = {}
        self._asa = {}

    def _rank_table(self, n, ratios, device):
        key = (n, tuple(ratios.items()) if ratios is not None else None, str(device))
        if key not in self._rank:
            self._rank[key] = torch.from_numpy(rank_levels(n, ratios)).to(device)
        return self._rank[key]

    def asa_engine(self, **kw) -> AsaEngine:
        key = tuple(sorted(kw.items()))
        if key not in self._asa:
            self._asa[key] = AsaEngine(AsaKnobs.cog(estimator="sampled_max", **kw))
        return self._asa[key]

    @_on_tensor_device
    def level_mask(self, scores: torch.Tensor, ratios=None, force: int = 2):
        """transfer_attn_to_mask (N:154-207) on fp32 scores [B,H,nq,nk] -> (level mask u8, idx, cnt4)."""
        assert scores.is_cuda and scores.dtype == torch.float32
        scores = scores.contiguous()
        B, H, nq, nk = scores.shape
        table = self._rank_table(nk, ratios, scores.device)
        mask = torch.empty(B, H, nq, nk, dtype=torch.uint8, device=scores.device)
        idx = torch.empty(B, H, nq, nk, dtype=torch.int32, device=scores.device)
        cnt4 = torch.empty(B, H, nq, 4, dtype=torch.int32, device=scores.device)
        check(self.lib.blade_multilevel_mask(scores.data_ptr(), B, H, nq, nk, table.data_ptr(), int(force),
                                             mask.data_ptr(), idx.data_ptr(), cnt4.data_ptr(), current_stream()))
        return mask, idx, cnt4

    @_on_tensor_device
    def mask_to_index(self, level_mask: torch.Tensor):
        m = level_mask.to(torch.uint8).contiguous()
        B, H, nq, nk = m.shape
        idx = torch.empty(B, H, nq, nk, dtype=torch.int32, device=m.device)
        cnt4 = torch.empty(B, H, nq, 4, dtype=torch.int32, device=m.device)
        check(self.lib.blade_level_mask_to_index(m.data_ptr(), B, H, nq, nk, idx.data_ptr(), cnt4.data_ptr(),
                                                 current_stream()))
        return idx, cnt4

    @_on_tensor_device
    def pyramid(self, k: torch.Tensor, v: torch.Tensor):
        """K/V mean-pooled by 2, 4, 8 over the block-padded sequence (K9:1307-1316): [(k2,v2),(k4,v4),(k8,v8)]."""
        B, H, S, D = k.shape
        nb = -(-S // 128)
        outs = []
        for rows in (64, 32, 16):
            outs.append((torch.empty(B, H, nb * rows, D, dtype=k.dtype, device=k.device),
                         torch.empty(B, H, nb * rows, D, dtype=k.dtype, device=k.device)))
        check(self.lib.blade_multilevel_pyramid(C.byref(tensor_desc(k)), C.byref(tensor_desc(v)),
                                                outs[0][0].data_ptr(), outs[0][1].data_ptr(), outs[1][0].data_ptr(),
                                                outs[1][1].data_ptr(), outs[2][0].data_ptr(), outs[2][1].data_ptr(),
                                                current_stream()))
        return outs

    @_on_tensor_device
    def attention(self, q, k, v, pyr, idx, cnt4, dst_row=None, sm_scale=None, want_lse=False):
        """_fwd_kernel (K9:339-692): out [B,H,S,D] (a transposed view of [B,S,H,D] memory)."""
        B, H, S, D = q.shape
        out = torch.empty(B, S, H, D, dtype=q.dtype, device=q.device).transpose(1, 2)
        lse = torch.empty(B, H, S, dtype=torch.float32, device=q.device) if want_lse else None
        eng = self.asa_engine()
        ws = eng._park(q.device, D)
        scale = (1.0 / math.sqrt(D)) if sm_scale is None else float(sm_scale)
        descs = [tensor_desc(t) for pair in pyr for t in pair]                 # k2, v2, k4, v4, k8, v8
        check(self.lib.blade_multilevel_attn_fwd(
            C.byref(tensor_desc(q)), C.byref(tensor_desc(k)), C.byref(tensor_desc(v)), *[C.byref(d) for d in descs],
            idx.data_ptr(), cnt4.data_ptr(), idx.shape[-1], C.byref(tensor_desc(out)), ptr(lse), ptr(dst_row), scale,
            ws.data_ptr(), ws.numel(), current_stream()))
        return (out, lse) if want_lse else out

    @_on_tensor_device
    def attention_bwd(self, q, k, v, pyr, idx, cnt4, out, lse, d_out, sm_scale=None):
        """Backward of `attention` (replaces the Triton backward kernels K9:695-1237): (dq, dk, dv) in q.dtype."""
        B, H, S, D = q.shape
        Sk = k.shape[2]
        dq = torch.empty(B, S, H, D, dtype=q.dtype, device=q.device).transpose(1, 2)
        dk = torch.empty(B, Sk, H, D, dtype=q.dtype, device=q.device).transpose(1, 2)
        dv = torch.empty(B, Sk, H, D, dtype=q.dtype, device=q.device).transpose(1, 2)
        nbytes = int(self.lib.blade_multilevel_bwd_workspace_bytes(B, H, S, Sk, D))
        ws = self.asa_engine().workspace(q.device, nbytes)
        scale = (1.0 / math.sqrt(D)) if sm_scale is None else float(sm_scale)
        if d_out.stride(-1) != 1:
            d_out = d_out.contiguous()
        descs = [tensor_desc(t) for pair in pyr for t in pair]
        check(self.lib.blade_multilevel_attn_bwd(
            C.byref(tensor_desc(q)), C.byref(tensor_desc(k)), C.byref(tensor_desc(v)), *[C.byref(d) for d in descs],
            idx.data_ptr(), cnt4.data_ptr(), idx.shape[-1], C.byref(tensor_desc(out)), C.byref(tensor_desc(d_out)),
            lse.data_ptr(), scale, C.byref(tensor_desc(dq)), C.byref(tensor_desc(dk)), C.byref(tensor_desc(dv)),
            ws.data_ptr(), ws.numel(), current_stream()))
        return dq, dk, dv

    def forward(self, q, k, v, grid, text_len, ratios, rearrange=True, sample_offsets=None, return_debug=False):
        """AdaptiveBlockSparseAttnTrain.forward (N:237-267) as kernels: gather into curve order (text to the tail) ->
        sampled-max block scores (N:64-90) -> level mask (N:154-207) -> pyramid -> attention with the inverse
        permutation folded into the output store."""
        B, H, S, D = q.shape
        eng = self.asa_engine(width=grid[0], height=grid[1], depth=grid[2], text_length=text_len, use_rearrange=rearrange)
        (qr, kr, vr), _, _ = eng.prep(q, k, v, rearrange=rearrange, want_means=False, want_pool=False)
        if qr is None:
            qr, kr, vr = q, k, v
        qo, ko = sample_offsets if sample_offsets is not None else (eng.draw_offsets(B, H, q.device),
                                                                     eng.draw_offsets(B, H, q.device))
        scores = eng.scores_sampled(qr, kr, qo, ko)
        mask, idx, cnt4 = self.level_mask(scores, ratios, force_last)
        pyr = self.pyramid(kr, vr)
        out = self.attention(qr, kr, vr, pyr, idx, cnt4, dst_row=eng.src_row(q.device, S) if rearrange else None)
        if return_debug:
            return out, dict(scores=scores, mask=mask, idx=idx, cnt4=cnt4)
        return out


_ENGINE: Optional[MultiLevelEngine] = None


def _engine() -> MultiLevelEngine:
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = MultiLevelEngine()
    return _ENGINE


def transfer_attn_to_mask(attn, mask_ratios=None):
    """N:154-207: [B,H,seq,seq] block scores -> int32 level mask (0 skip, 1 full, 2/4/8 pooled)."""
    mask, _, _ = _engine().level_mask(attn.float(), mask_ratios, force_last)
    return mask.to(torch.int32)


class _MultiLevelAttention(torch.autograd.Function):
    """The autograd wrapper the reference builds around its Triton kernels (K9:1375-1611): forward = _fwd_kernel,
    backward = the dq / dk / dv kernels, both replaced by CUDA kernels behind the C ABI."""

    @staticmethod
    def forward(ctx, q, k, v, mask, sm_scale):
        e = _engine()
        qd, kd, vd = q.detach(), k.detach(), v.detach()
        idx, cnt4 = e.mask_to_index(mask)
        pyr = e.pyramid(kd, vd)
        out, lse = e.attention(qd, kd, vd, pyr, idx, cnt4, sm_scale=sm_scale, want_lse=True)
        ctx.save_for_backward(qd, kd, vd, out, lse, idx, cnt4, *[t for pair in pyr for t in pair])
        ctx.sm_scale = sm_scale
        return out

    @staticmethod
    def backward(ctx, d_out):
        qd, kd, vd, out, lse, idx, cnt4, *flat = ctx.saved_tensors
        pyr = [(flat[0], flat[1]), (flat[2], flat[3]), (flat[4], flat[5])]
        dq, dk, dv = _engine().attention_bwd(qd, kd, vd, pyr, idx, cnt4, out, lse, d_out.to(qd.dtype), sm_scale=ctx.sm_scale)
        return dq, dk, dv, None, None


def sparse_attention_fn(q, k, v, mask, sm_scale=None):
    """sparse_attention_factory(BLOCK_M=128, BLOCK_N=128) (N:9; K9:1578-1611): multi-level attention on a caller's
    level mask [B,H,ceil(S/128),ceil(S/128)].  Differentiable in q, k, v like the reference's autograd function."""
    if torch.is_grad_enabled() and any(t.requires_grad for t in (q, k, v)):
        return _MultiLevelAttention.apply(q, k, v, mask, sm_scale)
    e = _engine()
    idx, cnt4 = e.mask_to_index(mask)
    return e.attention(q, k, v, e.pyramid(k, v), idx, cnt4, sm_scale=sm_scale)


def adaptive_block_sparse_attn(q, k, v):
    """N:210-235: q,k,v already in curve order.  Returns (out, sparsity) with the reference's ratio-table statistic."""
    require_no_grad(q, k, v)
    m = sys.modules[__name__]
    out = _engine().forward(q, k, v, (m.width, m.height, m.depth), m.text_length, m.mask_ratios, rearrange=False)
    density = sum((b - a) / lv for lv, (a, b) in m.mask_ratios.items() if lv != 0)          # N:230-233
    return out, 1 - density


class AdaptiveBlockSparseAttnTrain(nn.Module):
    """N:237-267: `inner_attention(q, k, v) -> out`, all [B,H,S,D].  FORWARD ONLY."""

    def __init__(self):
        super().__init__()
        self.sparsity_acc = 0.0
        self.sparsity_counter = 0
        self.use_rearrange = sys.modules[__name__].use_rearrange

    def forward(self, q, k, v):
        require_no_grad(q, k, v)
        m = sys.modules[__name__]
        out = _engine().forward(q, k, v, (m.width, m.height, m.depth), m.text_length, m.mask_ratios,
                                rearrange=bool(self.use_rearrange))
        self.sparsity_acc += 1 - sum((b - a) / lv for lv, (a, b) in m.mask_ratios.items() if lv != 0)
        self.sparsity_counter += 1
        if self.sparsity_counter % 600 == 0:                                    # N:254-256
            print(f"sparsity: {self.sparsity_acc / self.sparsity_counter}")
        return out
