"""Drop-in mirror of the reference module
cogvideox/train/special_attentions_local/TrainRelated/cogvideo_blocksparseattn.py (C) -- same knobs and names
as the reference, B200 kernels underneath.  Differences from the wan flavour, as in the reference:
text tokens (first `text_length` rows) move to the sequence tail inside the Gilbert reorder (C:141-161),
retain bounds come from an fp32 tensor multiply (C:230-231), the last two block rows/cols are forced on
(C:247-248), defaults C:9-16.  See wanx_blocksparseattn.py in this package for the function map.
"""
from __future__ import annotations

import sys

import torch
import torch.nn as nn

from .asa import AsaEngine, AsaKnobs, gilbert_tables

# ----------------------------- parameters (C:9-16) -----------------------------
use_rearrange = True
max_retain_ratio = 0.1
min_retain_ratio = 0.05
width = 45
height = 30
depth = 13
sample_gap = 15
text_length = 226
# ---------------------- literals of the reference (W:62,325,341) ----------------
block_size = 128
num_keep = 32
energy_threshold = 0.95
estimator = "meanpool"   # north-star kernel (a); the reference's sampled-max estimator is "sampled_max"
exact_merge = True
_FLAVOR = "cog"

_engines = {}


def _knobs(**override) -> AsaKnobs:
    m = sys.modules[__name__]
    kw = dict(flavor=m._FLAVOR, use_rearrange=m.use_rearrange, max_retain_ratio=m.max_retain_ratio,
              min_retain_ratio=m.min_retain_ratio, width=m.width, height=m.height, depth=m.depth,
              sample_gap=m.sample_gap, text_length=m.text_length, block_size=m.block_size, num_keep=m.num_keep,
              energy_threshold=m.energy_threshold, estimator=m.estimator, exact_merge=m.exact_merge)
    kw.update(override)
    return AsaKnobs(**kw)


def _engine(**override) -> AsaEngine:
    kn = _knobs(**override)
    key = tuple(sorted(kn.__dict__.items()))
    if key not in _engines:
        _engines[key] = AsaEngine(kn)
    return _engines[key]


def simple_pooling(x, sample_gap=None):
    """W:88-93."""
    m = sys.modules[__name__]
    eng = _engine(sample_gap=sample_gap or m.sample_gap, use_rearrange=False)
    _, _, (kp, _vp) = eng.prep(x, x, x, rearrange=False, want_means=False, want_pool=True)
    return kp


def transfer_attn_to_mask(attn, mode="energy", init_k=None, max_retain_ratio=0.7, min_retain_ratio=0.1,
                          energy_threshold=0.95):
    """W:162-233: [B,H,nb,nb] block scores -> bool mask.  Only mode="energy" is live upstream (W:337)."""
    if mode == "topk":
        if init_k is None:
            raise ValueError("init_k is required in topk mode")          # W:193-194
        raise ValueError("mode 'topk' is dead code in the reference (W:337 always passes 'energy')")
    if mode != "energy":
        raise ValueError(f"unsupported mode: {mode}")                      # W:232
    import numpy as np
    seq = attn.shape[2]
    eng = _engine()

    def _bound(r):                       # C:230-231: (seq * ratio_tensor).to(int), clamp(min=1)
        if torch.is_tensor(r):
            vals = torch.clamp((seq * r.float()).to(torch.int), min=1).flatten().tolist()
            if len(set(vals)) != 1:
                raise ValueError("per-head retain bounds must agree in this entry point")
            return vals[0]
        return max(1, int(np.float32(seq) * np.float32(r)))

    _, _, mask = eng.select(attn.float(), lo=_bound(min_retain_ratio), hi=_bound(max_retain_ratio), force_last=2,
                            thr=energy_threshold)                                   # C:247-248 forced rows/cols
    return mask


def block_sparse_attn(q, k, v, block_mask):
    """W:278-309: returns (out [B,H,S,D], lse [B,H,S,1] in q.dtype)."""
    assert q.shape == k.shape == v.shape                                   # W:250-251
    eng = _engine()
    nq = -(-q.size(2) // 128)
    nk = -(-k.size(2) // 128)
    idx, cnt = eng.mask_to_index(block_mask[:, :, :nq, :nk])               # crop the S//128+1 quirk (W:22)
    out, lse = eng.block_sparse_attn(q, k, v, idx, cnt)
    return out, lse.unsqueeze(-1).to(q.dtype)


def standard_attn(q, k, v):
    """W:21-24: dense attention through the same kernel with an all-ones block mask."""
    eng = _engine()
    nq = -(-q.size(2) // 128)
    nk = -(-k.size(2) // 128)
    ones = torch.ones(q.size(0), q.size(1), nq, nk, dtype=torch.bool, device=q.device)
    idx, cnt = eng.mask_to_index(ones)
    out, lse = eng.block_sparse_attn(q, k, v, idx, cnt)
    return out, lse.unsqueeze(-1).to(q.dtype)


def adaptive_block_sparse_attn(q, k, v):
    """W:311-372: q,k,v already in Gilbert order.  Returns (out, sparsity) with sparsity a 0-dim DEVICE
    tensor (the reference's float statistic W:372 without its host sync)."""
    eng = _engine(use_rearrange=False)
    out, cnt = eng.forward(q, k, v)
    nb = cnt.shape[-1]
    m = sys.modules[__name__]
    sparsity = 1 - cnt.sum().float() / float(cnt.numel() * nb) - 1.0 / m.sample_gap
    return out, sparsity


class GilbertRearranger(nn.Module):
    """W:102-159 -- kept for API parity; the engine fuses these gathers into its kernels."""

    def __init__(self, width, height, depth, text_length=224):
        super().__init__()
        self.width, self.height, self.depth = width, height, depth
        self.total_elements = width * height * depth
        self.text_length = text_length
        c2r, r2c = gilbert_tables(width, height, depth)
        self.register_buffer("original_order2gilbert_order", torch.from_numpy(c2r))
        self.register_buffer("gilbert_order2original_order", torch.from_numpy(r2c))

    def rearrange(self, q, k, v):                                  # C:141-154
        o, t = self.original_order2gilbert_order, self.text_length

        def one(x):
            return torch.cat((x[..., t:, :].index_select(-2, o), x[..., :t, :]), dim=-2)
        return one(q), one(k), one(v)

    def reversed_rearrange(self, out):                             # C:156-161
        t = self.text_length
        vid, txt = out[..., :-t, :], out[..., -t:, :]
        return torch.cat((txt, vid.index_select(-2, self.gilbert_order2original_order)), dim=-2)


class AdaptiveBlockSparseAttnTrain(nn.Module):
    """W:375-408: `inner_attention(q, k, v) -> out`, all [B,H,S,D]."""

    def __init__(self):
        super().__init__()
        m = sys.modules[__name__]
        self.gilbert_rearranger = GilbertRearranger(m.width, m.height, m.depth, m.text_length)
        self.sparsity_acc = 0.0
        self.sparsity_counter = 0
        self.use_rearrange = m.use_rearrange
        self._cnt_acc = None
        self._cnt_den = 0
        self.print_every = 800

    supports_fused_rope = True
    supports_fused_qk_norm = True  # per-head LayerNorm of q/k (MC:54-57) inside the gather kernel   # processors may hand over un-rotated q/k plus the rotary table

    def forward(self, q, k, v, virtual_rows=None, rotary=None, qk_norm=None):
        """`virtual_rows` (optional, int32 [S]): q/k/v are strided views into a packed Ulysses receive buffer and
        token s lives at row virtual_rows[s] (video_blade_b200.ulysses.scatter_heads_fused).
        `rotary` (optional): (fp32 table [rows, D/2, 2] of (cos, sin), first_row) -- the processor's rotary
        embedding (MW:108-116 / MC:59-64) is then applied to q and k inside the gather kernel."""
        m = sys.modules[__name__]
        eng = _engine(use_rearrange=bool(self.use_rearrange))
        out, cnt = eng.forward(q, k, v, virtual_rows=virtual_rows, rope=rotary, qk_norm=qk_norm)
        # sparsity bookkeeping without the reference's per-layer .item() sync (W:398): accumulate on device
        if self._cnt_acc is None or self._cnt_acc.device != cnt.device:
            self._cnt_acc = torch.zeros((), dtype=torch.float64, device=cnt.device)
        self._cnt_acc += cnt.sum()
        self._cnt_den += cnt.numel() * cnt.shape[-1]
        self.sparsity_counter += 1
        if self.print_every and self.sparsity_counter % self.print_every == 0:
            print(f"sparsity: {self.average_sparsity()}")
        return out

    def average_sparsity(self) -> float:
        """Running mean of `1 - mask.mean() - 1/sample_gap` (W:372,401-403); syncs only when asked."""
        if self._cnt_acc is None:
            return 0.0
        m = sys.modules[__name__]
        self.sparsity_acc = float(self.sparsity_counter) * (
            1.0 - float(self._cnt_acc.item()) / max(1, self._cnt_den) - 1.0 / m.sample_gap)
        return self.sparsity_acc / max(1, self.sparsity_counter)
