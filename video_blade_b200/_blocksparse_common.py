"""Shared implementation behind the two drop-in mirrors `wanx_blocksparseattn.py` (W) and
`cogvideo_blocksparseattn.py` (C).  The reference ships two near-identical modules; here each mirror keeps only
its module-level knobs (read at call time, like the reference's globals) and `build_api(module_name)` supplies the
functions and classes, bound to that module's knobs.

    reference (W / C)                              here
    ---------------------------------------------  ------------------------------------------------
    use_rearrange ... text_length  (W:9-16/C:9-16)  identical names and defaults, read at call time
    standard_attn(q,k,v)               (W:21-24)    AsaEngine.block_sparse_attn with an all-ones list
    simple_pooling(x, sample_gap)      (W:88-93)    blade_asa_prep (pool kernel)
    GilbertRearranger          (W:102-159/C:110-161)  blade_gilbert_tables + gather fused into prep/attn
    transfer_attn_to_mask(...) (W:162-233/C:177-249)  blade_asa_select ("energy"; "topk" is dead code upstream)
    block_sparse_attn(q,k,v,mask)      (W:278-309)  blade_mask_to_index + blade_block_sparse_attn_fwd
    adaptive_block_sparse_attn (W:311-372/C:327-394)  blade_asa_forward (use_rearrange False)
    AdaptiveBlockSparseAttnTrain (W:375-408/C:398-427)  blade_asa_forward

Forward only: the C ABI has no backward entry point, so every entry point raises when autograd would need one
(`asa.require_no_grad`) instead of silently returning an output without grad_fn.
"""
from __future__ import annotations

import sys

import numpy as np
import torch
import torch.nn as nn

from .asa import AsaEngine, AsaKnobs, gilbert_tables, require_no_grad

_engines = {}


def build_api(module_name: str) -> dict:
    def mod():
        return sys.modules[module_name]

    def _knobs(**override) -> AsaKnobs:
        m = mod()
        kw = dict(flavor=m._FLAVOR, use_rearrange=m.use_rearrange, max_retain_ratio=m.max_retain_ratio,
                  min_retain_ratio=m.min_retain_ratio, width=m.width, height=m.height, depth=m.depth,
                  sample_gap=m.sample_gap, text_length=m.text_length, block_size=m.block_size, num_keep=m.num_keep,
                  energy_threshold=m.energy_threshold, estimator=m.estimator, exact_merge=m.exact_merge,
                  select_rounding=getattr(m, "select_rounding", "fp32"))
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
        eng = _engine(sample_gap=sample_gap or mod().sample_gap, use_rearrange=False)
        _, _, (kp, _vp) = eng.prep(x, x, x, rearrange=False, want_means=False, want_pool=True)
        return kp

    def transfer_attn_to_mask(attn, mode="energy", init_k=None, max_retain_ratio=0.7, min_retain_ratio=0.1,
                              energy_threshold=0.95):
        """W:162-233 / C:177-249: [B,H,nb,nb] block scores -> bool mask.  Only mode="energy" is live upstream (W:337)."""
        if mode == "topk":
            if init_k is None:
                raise ValueError("init_k is required in topk mode")          # W:193-194
            raise ValueError("mode 'topk' is dead code in the reference (W:337 always passes 'energy')")
        if mode != "energy":
            raise ValueError(f"unsupported mode: {mode}")                      # W:232
        cog = mod()._FLAVOR == "cog"
        seq = attn.shape[2] if cog else attn.shape[-1]

        def _bound(r):
            if torch.is_tensor(r):                   # C:230-231: (seq * ratio_tensor).to(int), clamp(min=1)
                vals = torch.clamp((seq * r.float()).to(torch.int), min=1).flatten().tolist()
                if len(set(vals)) != 1:
                    raise ValueError("per-head retain bounds must agree in this entry point")
                return vals[0]
            if cog:
                return max(1, int(np.float32(seq) * np.float32(r)))
            return max(1, int(seq * r))              # W:215-216

        _, _, mask = _engine().select(attn.float(), lo=_bound(min_retain_ratio), hi=_bound(max_retain_ratio),
                                      force_last=2 if cog else 0, thr=energy_threshold)   # C:247-248 forced rows/cols
        return mask

    def _dense_or_masked(q, k, v, block_mask):
        require_no_grad(q, k, v)
        eng = _engine()
        nq = -(-q.size(2) // 128)
        nk = -(-k.size(2) // 128)
        if block_mask is None:
            block_mask = torch.ones(q.size(0), q.size(1), nq, nk, dtype=torch.bool, device=q.device)
        idx, cnt = eng.mask_to_index(block_mask[:, :, :nq, :nk])           # crop the S//128+1 quirk (W:22)
        out, lse = eng.block_sparse_attn(q, k, v, idx, cnt)
        return out, lse.unsqueeze(-1).to(q.dtype)

    def block_sparse_attn(q, k, v, block_mask):
        """W:278-309: returns (out [B,H,S,D], lse [B,H,S,1] in q.dtype)."""
        assert q.shape == k.shape == v.shape                                   # W:250-251
        return _dense_or_masked(q, k, v, block_mask)

    def standard_attn(q, k, v):
        """W:21-24: dense attention through the same kernel with an all-ones block mask."""
        return _dense_or_masked(q, k, v, None)

    def adaptive_block_sparse_attn(q, k, v):
        """W:311-372: q,k,v already in Gilbert order.  Returns (out, sparsity) with sparsity a 0-dim DEVICE
        tensor (the reference's float statistic W:372 without its host sync)."""
        require_no_grad(q, k, v)
        out, cnt = _engine(use_rearrange=False).forward(q, k, v)
        nb = cnt.shape[-1]
        sparsity = 1 - cnt.sum().float() / float(cnt.numel() * nb) - 1.0 / mod().sample_gap
        return out, sparsity

    class GilbertRearranger(nn.Module):
        """W:102-159 / C:110-161 -- kept for API parity; the engine fuses these gathers into its kernels."""

        def __init__(self, width, height, depth, text_length=224):
            super().__init__()
            self.width, self.height, self.depth = width, height, depth
            self.total_elements = width * height * depth
            self.text_length = text_length
            c2r, r2c = gilbert_tables(width, height, depth)
            self.register_buffer("original_order2gilbert_order", torch.from_numpy(c2r))
            self.register_buffer("gilbert_order2original_order", torch.from_numpy(r2c))
            self._text_to_tail = mod()._FLAVOR == "cog"

        def rearrange(self, q, k, v):
            o, t = self.original_order2gilbert_order, self.text_length
            if not self._text_to_tail:
                return q.index_select(-2, o), k.index_select(-2, o), v.index_select(-2, o)

            def one(x):                                                    # C:141-154
                return torch.cat((x[..., t:, :].index_select(-2, o), x[..., :t, :]), dim=-2)
            return one(q), one(k), one(v)

        def reversed_rearrange(self, out):
            if not self._text_to_tail:
                return out.index_select(-2, self.gilbert_order2original_order)
            t = self.text_length                                           # C:156-161
            vid, txt = out[..., :-t, :], out[..., -t:, :]
            return torch.cat((txt, vid.index_select(-2, self.gilbert_order2original_order)), dim=-2)

    class AdaptiveBlockSparseAttnTrain(nn.Module):
        """W:375-408 / C:398-427: `inner_attention(q, k, v) -> out`, all [B,H,S,D].  FORWARD ONLY (inference): raises
        if q/k/v or the fused norm weights require grad while autograd is recording."""

        def __init__(self):
            super().__init__()
            m = mod()
            self.gilbert_rearranger = GilbertRearranger(m.width, m.height, m.depth, m.text_length)
            self.sparsity_acc = 0.0
            self.sparsity_counter = 0
            self.use_rearrange = m.use_rearrange
            self._cnt_acc = None
            self._cnt_den = 0
            self.print_every = 800 if m._FLAVOR == "cog" else 200

        # What the processors may hand over instead of doing it in torch.  Every engine path (one-call mean-pool /
        # sampled-max, staged block-64) applies rope and norm inside the gather kernel, so these hold for all knobs.
        supports_fused_rope = True     # un-rotated q/k plus the rotary table (MW:108-116 / MC:59-64)
        supports_fused_qk_norm = True  # un-normalised q/k plus the norm weights (MW:99-102 RMSNorm / MC:54-57 LayerNorm)

        def forward(self, q, k, v, virtual_rows=None, rotary=None, qk_norm=None, **engine_kw):
            """`virtual_rows` (optional, int32 [S]): q/k/v are strided views into a packed Ulysses receive buffer and
            token s lives at row virtual_rows[s] (video_blade_b200.ulysses.scatter_heads_fused).
            `rotary` (optional): (fp32 table [rows, D/2, 2] of (cos, sin), first_row) -- the processor's rotary
            embedding (MW:108-116 / MC:59-64) is then applied to q and k inside the gather kernel."""
            require_no_grad(q, k, v, *((qk_norm[1], qk_norm[2]) if qk_norm is not None else ()))
            eng = _engine(use_rearrange=bool(self.use_rearrange))
            out, cnt = eng.forward(q, k, v, virtual_rows=virtual_rows, rope=rotary, qk_norm=qk_norm,
                                   selected_acc=self.counter(q.device), **engine_kw)
            self.count_call(cnt)
            return out

        def counter(self, device):
            """Sparsity bookkeeping without the reference's per-layer .item() sync (W:398) and without extra launches:
            the selection kernel adds its count of selected block pairs to this device counter."""
            if self._cnt_acc is None or self._cnt_acc.device != device:
                self._cnt_acc = torch.zeros(1, dtype=torch.int64, device=device)
            return self._cnt_acc

        def count_call(self, cnt):
            self._cnt_den += cnt.numel() * cnt.shape[-1]
            self.sparsity_counter += 1
            if self.print_every and self.sparsity_counter % self.print_every == 0:
                print(f"sparsity: {self.average_sparsity()}")

        def average_sparsity(self) -> float:
            """Running mean of `1 - mask.mean() - 1/sample_gap` (W:372,401-403); syncs only when asked."""
            if self._cnt_acc is None:
                return 0.0
            self.sparsity_acc = float(self.sparsity_counter) * (
                1.0 - float(self._cnt_acc.item()) / max(1, self._cnt_den) - 1.0 / mod().sample_gap)
            return self.sparsity_acc / max(1, self.sparsity_counter)

    for cls in (GilbertRearranger, AdaptiveBlockSparseAttnTrain):
        cls.__module__ = module_name
        cls.__qualname__ = cls.__name__
    return dict(_knobs=_knobs, _engine=_engine, _engines=_engines, simple_pooling=simple_pooling,
                transfer_attn_to_mask=transfer_attn_to_mask, block_sparse_attn=block_sparse_attn,
                standard_attn=standard_attn, adaptive_block_sparse_attn=adaptive_block_sparse_attn,
                GilbertRearranger=GilbertRearranger, AdaptiveBlockSparseAttnTrain=AdaptiveBlockSparseAttnTrain)
