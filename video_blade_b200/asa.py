"""Host side of the ASA hot path: knobs, permutation tables, workspace and the C-ABI calls.

Mirrors what the reference module does around its kernels
(wanx_blocksparseattn.py W:311-408 / cogvideo_blocksparseattn.py C:327-427) but every tensor op of the
reference (index_select, pad, topk, sort, cumsum, scatter, the two external attention calls, the merge)
is one of the hand-written kernels behind include/blade_asa.h.  PyTorch is used for device memory and
streams only.
"""
from __future__ import annotations

import ctypes as C
import functools
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BladeQkNorm, BladeAsaConfig, check, current_stream, ptr, tensor_desc


@dataclass
class AsaKnobs:
    """The reference's module-level parameters (W:9-16 / C:9-16) plus its literals (W:62,325,341)."""
    flavor: str = "wan"
    use_rearrange: bool = True
    max_retain_ratio: float = 0.17
    min_retain_ratio: float = 0.05
    width: int = 52
    height: int = 30
    depth: int = 21
    sample_gap: int = 30
    text_length: int = 0
    block_size: int = 128
    num_keep: int = 32
    energy_threshold: float = 0.95
    estimator: str = "meanpool"     # "meanpool" (north-star kernel (a)); "sampled_max" = reference P (next)
    exact_merge: bool = True        # reproduce the bf16 op chain of W:351-370
    select_rounding: str = "fp32"   # "bf16"/"f16": prefix sums and threshold rounded like torch on a half-precision Po

    @staticmethod
    def wan(**kw) -> "AsaKnobs":
        return AsaKnobs(flavor="wan", **kw)

    @staticmethod
    def cog(**kw) -> "AsaKnobs":
        base = dict(flavor="cog", max_retain_ratio=0.1, width=45, height=30, depth=13, sample_gap=15,
                    text_length=226)
        base.update(kw)
        return AsaKnobs(**base)

    def retain_bounds(self, nb: int) -> Tuple[int, int]:
        """W:215-216 (Python double arithmetic) / C:230-231 (fp32 tensor multiply, truncated)."""
        if self.flavor == "cog":
            lo = int(np.float32(nb) * np.float32(self.min_retain_ratio))
            hi = int(np.float32(nb) * np.float32(self.max_retain_ratio))
        else:
            lo = int(nb * self.min_retain_ratio)
            hi = int(nb * self.max_retain_ratio)
        return max(1, lo), max(1, hi)

    def c_config(self, nb: int) -> BladeAsaConfig:
        lo, hi = self.retain_bounds(nb)
        cfg = BladeAsaConfig()
        cfg.block_size = self.block_size
        cfg.sample_gap = self.sample_gap
        cfg.min_retain = lo
        cfg.max_retain = hi
        cfg.energy_threshold = self.energy_threshold
        cfg.force_last = 2 if self.flavor == "cog" else 0
        cfg.num_keep = self.num_keep
        cfg.estimator = 0 if self.estimator == "meanpool" else 1
        cfg.exact_merge = 1 if self.exact_merge else 0
        cfg.select_rounding = {"fp32": 0, "bf16": 1, "f16": 2}[self.select_rounding]
        return cfg


def gilbert_tables(width: int, height: int, depth: int) -> Tuple[np.ndarray, np.ndarray]:
    """(curve2raster, raster2curve) int64 arrays from the library's host implementation (W:102-129)."""
    n = width * height * depth
    c2r = np.empty(n, np.int64)
    r2c = np.empty(n, np.int64)
    check(_lib.load().blade_gilbert_tables(width, height, depth, c2r.ctypes.data, r2c.ctypes.data))
    return c2r, r2c


def token_order(knobs: AsaKnobs) -> np.ndarray:
    """src_row[r] = index (in the caller's token order) of the token that sits at Gilbert-order row r.
    wan: curve2raster (W:146-148).  cog: video tokens in curve order, then the text tokens that came first
    (C:144-154).  The inverse move (W:154-159 / C:156-161) is a scatter through the same table."""
    c2r, _ = gilbert_tables(knobs.width, knobs.height, knobs.depth)
    if knobs.text_length:
        t = knobs.text_length
        return np.concatenate([c2r + t, np.arange(t, dtype=np.int64)]).astype(np.int32)
    return c2r.astype(np.int32)


def _on_tensor_device(fn):
    """Every C-ABI launch goes to the CURRENT device's stream (`_lib.current_stream`) and the library reads
    cudaGetDevice for its per-device state: run the entry point with the tensors' device current, and refuse
    tensors that live on different devices (a foreign pointer on the wrong stream is an illegal access at best)."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kw):
        dev = None
        for a in list(args) + list(kw.values()):
            for t in (a if isinstance(a, (tuple, list)) else (a,)):
                if torch.is_tensor(t) and t.is_cuda:
                    if dev is None:
                        dev = t.device
                    elif t.device != dev:
                        raise RuntimeError(f"video_blade_b200: tensors on different devices ({dev} and {t.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(self, *args, **kw)
        with torch.cuda.device(dev):
            return fn(self, *args, **kw)
    return wrapper


def require_no_grad(*tensors):
    """The C ABI is forward-only (no backward entry point): an output written through raw pointers has no grad_fn, so
    a training caller (the reference module sits inside train_*_tdm.py) would silently get zero gradient through the
    attention.  Fail loudly instead."""
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors):
        raise RuntimeError(
            "video_blade_b200 ASA is forward-only: q/k/v (or the fused norm weights) require grad but the CUDA path has "
            "no backward kernel.  Run under torch.no_grad() / detach the inputs (inference), or use the reference "
            "module for training.")


class AsaEngine:
    """Owns device-side tables + workspace for one (device, knobs) and drives the kernels."""

    def __init__(self, knobs: AsaKnobs):
        self.knobs = knobs
        self.lib = _lib.load()
        self._src_row = {}      # device -> int32 tensor
        self._ws = {}           # (device, bytes) -> uint8 tensor
        self._sel_acc = {}      # device -> int64 [1] running count of selected block pairs
        self._order_np: Optional[np.ndarray] = None

    # ---- tables / workspace --------------------------------------------------------------
    def src_row(self, device, S: int) -> Optional[torch.Tensor]:
        if not self.knobs.use_rearrange:
            return None
        key = (str(device), S)
        if key not in self._src_row:
            if self._order_np is None:
                self._order_np = token_order(self.knobs)
            if self._order_np.size != S:
                raise ValueError(
                    f"sequence length {S} != width*height*depth + text_length = {self._order_np.size} "
                    f"({self.knobs.width}x{self.knobs.height}x{self.knobs.depth}+{self.knobs.text_length})")
            self._src_row[key] = torch.from_numpy(self._order_np).to(device)
        return self._src_row[key]

    def workspace(self, device, nbytes: int) -> torch.Tensor:
        """Scratch for one layer call (gathered copies, means, lists, the attention item counter and parked tiles).
        One buffer per (device, CUDA stream): calls on the same stream are ordered and may share it; calls in flight
        on different streams (CFG branches, worker threads) must not, and do not."""
        key = (str(device), torch.cuda.current_stream(device).cuda_stream)
        cur = self._ws.get(key)
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            self._ws[key] = cur
        off = (-cur.data_ptr()) % 1024
        return cur[off:off + nbytes]

    @staticmethod
    def _require_cuda(*ts):
        for t in ts:
            if t is not None and not getattr(t, "is_cuda", False):
                raise RuntimeError("video_blade_b200 runs on CUDA tensors only (no CPU fallback)")

    # ---- stage entry points (same roles as the reference helpers) ------------------------
    @_on_tensor_device
    def select(self, scores: torch.Tensor, lo=None, hi=None, force_last=None, thr=None,
               want_mask=True):
        """transfer_attn_to_mask(mode='energy') on fp32 scores [B,H,nq,nk] -> (idx, cnt, mask)."""
        self._require_cuda(scores)
        assert scores.dtype == torch.float32
        scores = scores.contiguous()
        B, H, nq, nk = scores.shape
        cfg = self.knobs.c_config(nk)
        if lo is not None:
            cfg.min_retain = int(lo)
        if hi is not None:
            cfg.max_retain = int(hi)
        if force_last is not None:
            cfg.force_last = int(force_last)
        if thr is not None:
            cfg.energy_threshold = float(thr)
        idx = torch.empty(B, H, nq, nk, dtype=torch.int32, device=scores.device)
        cnt = torch.empty(B, H, nq, dtype=torch.int32, device=scores.device)
        mask = torch.empty(B, H, nq, nk, dtype=torch.uint8, device=scores.device) if want_mask else None
        check(self.lib.blade_asa_select(scores.data_ptr(), B, H, nq, nk, C.byref(cfg), None, None,
                                        idx.data_ptr(), cnt.data_ptr(), ptr(mask), None, current_stream()))
        return idx, cnt, (mask.bool() if want_mask else None)

    @_on_tensor_device
    def mask_to_index(self, mask: torch.Tensor):
        self._require_cuda(mask)
        m = mask.to(torch.uint8).contiguous()
        B, H, nq, nk = m.shape
        idx = torch.empty(B, H, nq, nk, dtype=torch.int32, device=m.device)
        cnt = torch.empty(B, H, nq, dtype=torch.int32, device=m.device)
        check(self.lib.blade_mask_to_index(m.data_ptr(), B, H, nq, nk, idx.data_ptr(), cnt.data_ptr(),
                                           current_stream()))
        return idx, cnt

    @_on_tensor_device
    def prep(self, q, k, v, rearrange: bool, want_means=True, want_pool=True, rope=None):
        """Gather into Gilbert order (optional) + block means + gap-pooled K/V.  rope = (table fp32 [rows,D/2,2],
        first_row): rotary embedding of q and k fused into the gather (needs the output copies)."""
        self._require_cuda(q, k, v)
        B, H, S, D = q.shape
        kn = self.knobs
        dev = q.device
        src = self.src_row(dev, S) if rearrange else None
        nb = -(-S // kn.block_size)
        q_r = k_r = v_r = None
        if src is not None or rope is not None:
            q_r = torch.empty(B, H, S, D, dtype=q.dtype, device=dev)
            k_r = torch.empty_like(q_r)
            v_r = torch.empty_like(q_r)
        qm = km = kp = vp = None
        if want_means:
            qm = torch.empty(B, H, nb, D, dtype=torch.float32, device=dev)
            km = torch.empty_like(qm)
        gap = kn.sample_gap if want_pool else 0
        if gap:
            npool = -(-S // gap)
            kp = torch.empty(B, H, npool, D, dtype=q.dtype, device=dev)
            vp = torch.empty_like(kp)
        table, first = rope if rope is not None else (None, 0)
        check(self.lib.blade_asa_prep_rope(C.byref(tensor_desc(q)), C.byref(tensor_desc(k)), C.byref(tensor_desc(v)),
                                           ptr(src), ptr(q_r), ptr(k_r), ptr(v_r), ptr(qm), ptr(km), ptr(kp), ptr(vp),
                                           kn.block_size, gap, ptr(table), int(first), current_stream()))
        return (q_r, k_r, v_r), (qm, km), (kp, vp)

    @_on_tensor_device
    def scores_meanpool(self, qm: torch.Tensor, km: torch.Tensor) -> torch.Tensor:
        B, H, nb, D = qm.shape
        sc = torch.empty(B, H, nb, nb, dtype=torch.float32, device=qm.device)
        check(self.lib.blade_asa_scores_meanpool(qm.data_ptr(), km.data_ptr(), sc.data_ptr(), B, H, nb, D,
                                                 current_stream()))
        return sc

    # ---- the reference's sampled-max estimator (W:62-87 -> P) ------------------------------
    def draw_offsets(self, B, H, device, generator=None):
        """W:49-51: rand[B,H,1,block] -> topk indices; one set of num_keep offsets per (b,h)."""
        kn = self.knobs
        rand = torch.rand(B, H, 1, kn.block_size, device=device, generator=generator)
        return torch.topk(rand, kn.num_keep, dim=3).indices[:, :, 0].to(torch.int32).contiguous()

    @_on_tensor_device
    def scores_sampled(self, q, k, q_off, k_off):
        """efficient_attn_with_pooling (W:62-87): fp32 [B,H,nb,nb] holding the q.dtype-rounded Po."""
        self._require_cuda(q, k, q_off, k_off)
        B, H, S, D = q.shape
        kn = self.knobs
        assert kn.num_keep == 32, "the estimator kernel is built for num_keep = 32 (W:62)"
        nb = -(-S // kn.block_size)
        q_s = torch.empty(B, H, nb * 32, D, dtype=q.dtype, device=q.device)
        k_s = torch.empty_like(q_s)
        q_off = q_off.to(torch.int32).contiguous()
        k_off = k_off.to(torch.int32).contiguous()
        check(self.lib.blade_asa_sample_tokens(C.byref(tensor_desc(q)), C.byref(tensor_desc(k)), q_off.data_ptr(),
                                               k_off.data_ptr(), q_s.data_ptr(), k_s.data_ptr(), kn.block_size,
                                               current_stream()))
        sc = torch.empty(B, H, nb, nb, dtype=torch.float32, device=q.device)
        check(self.lib.blade_asa_scores_sampled(q_s.data_ptr(), k_s.data_ptr(), sc.data_ptr(), B, H, nb, D,
                                                _lib._dtype_code(q), current_stream()))
        return sc

    def _park(self, device, D):
        return self.workspace(device, int(self.lib.blade_attn_workspace_bytes(D)))

    @_on_tensor_device
    def block_sparse_attn(self, q, k, v, idx, cnt, out=None, dst_row=None, want_lse=True, sub64=False):
        """block_sparse_attn(q,k,v,block_mask) (W:278-309) on an index list; returns (out, lse fp32 [B,H,S])."""
        self._require_cuda(q, k, v, idx, cnt)
        B, H, S, D = q.shape
        if out is None:
            out = torch.empty(B, S, H, D, dtype=q.dtype, device=q.device).transpose(1, 2)
        lse = torch.empty(B, H, S, dtype=torch.float32, device=q.device) if want_lse else None
        ws = self._park(q.device, D)
        od = tensor_desc(out)
        fn = self.lib.blade_block_sparse_attn64_fwd if sub64 else self.lib.blade_block_sparse_attn_fwd
        check(fn(
            C.byref(tensor_desc(q)), C.byref(tensor_desc(k)), C.byref(tensor_desc(v)), idx.data_ptr(), cnt.data_ptr(),
            idx.shape[-1], C.byref(od), ptr(lse), ptr(dst_row), 1.0 / math.sqrt(D), ws.data_ptr(), ws.numel(),
            current_stream()))
        return out, lse

    @_on_tensor_device
    def mask64_to_index(self, mask64: torch.Tensor):
        """block_size 64: bool mask [B,H,nq64,nk64] -> quadrant-flagged list over 128x128 tiles."""
        self._require_cuda(mask64)
        m = mask64.to(torch.uint8).contiguous()
        B, H, nq64, nk64 = m.shape
        nq, nk = -(-nq64 // 2), -(-nk64 // 2)
        idx = torch.empty(B, H, nq, nk, dtype=torch.int32, device=m.device)
        cnt = torch.empty(B, H, nq, dtype=torch.int32, device=m.device)
        check(self.lib.blade_mask64_to_index(m.data_ptr(), B, H, nq64, nk64, idx.data_ptr(), cnt.data_ptr(),
                                             current_stream()))
        return idx, cnt

    @_on_tensor_device
    def asa_attn(self, q, k, v, idx, cnt, k_pool, v_pool, out=None, dst_row=None, exact_merge=None, sub64=False):
        """Sparse branch + pooled branch + merge (W:343-370) in one launch."""
        self._require_cuda(q, k, v, idx, cnt, k_pool, v_pool)
        B, H, S, D = q.shape
        if out is None:
            out = torch.empty(B, S, H, D, dtype=q.dtype, device=q.device).transpose(1, 2)
        ws = self._park(q.device, D)
        em = self.knobs.exact_merge if exact_merge is None else exact_merge
        fn = self.lib.blade_asa_attn64_fwd if sub64 else self.lib.blade_asa_attn_fwd
        check(fn(
            C.byref(tensor_desc(q)), C.byref(tensor_desc(k)), C.byref(tensor_desc(v)), idx.data_ptr(), cnt.data_ptr(),
            idx.shape[-1], C.byref(tensor_desc(k_pool)), C.byref(tensor_desc(v_pool)), self.knobs.sample_gap,
            C.byref(tensor_desc(out)), ptr(dst_row), 1.0 / math.sqrt(D), 1 if em else 0, ws.data_ptr(), ws.numel(),
            current_stream()))
        return out

    @_on_tensor_device
    def qk_rms_stat(self, q, k, eps: float):
        """rstd of the processor's RMSNorm over all heads' channels (MW:99-102) for q and k: fp32 [2, B*S].
        q, k: [B,H,S,D] views of token-major [B,S,H*D] memory."""
        self._require_cuda(q, k)
        B, H, S, D = q.shape
        out = torch.empty(2, B * S, dtype=torch.float32, device=q.device)
        check(self.lib.blade_qk_rms_stat(C.byref(tensor_desc(q)), C.byref(tensor_desc(k)), float(eps), out.data_ptr(),
                                         current_stream()))
        return out

    # ---- the whole layer ------------------------------------------------------------------
    def selected_counter(self, device) -> torch.Tensor:
        """Device counter (uint64 as int64 [1]) that every layer call adds its number of selected (q-block, k-block)
        pairs to -- the numerator of the reference's sparsity statistic (W:372) without the per-layer `.item()` sync
        (W:398) and without extra torch launches."""
        key = str(device)
        if key not in self._sel_acc:
            self._sel_acc[key] = torch.zeros(1, dtype=torch.int64, device=device)
        return self._sel_acc[key]

    @_on_tensor_device
    def forward(self, q, k, v, scores: Optional[torch.Tensor] = None, return_debug: bool = False,
                virtual_rows: Optional[torch.Tensor] = None, sample_offsets=None, rope=None, qk_norm=None,
                peers=None, out: Optional[torch.Tensor] = None, selected_acc: Optional[torch.Tensor] = None):
        """AdaptiveBlockSparseAttnTrain.forward (W:383-408 / C:405-427): q,k,v [B,H,S,D] in the caller's
        token order (strided views allowed) -> out [B,H,S,D] (a transposed view of [B,S,H,D] memory, so the
        processor's `.transpose(1,2).flatten(2,3)` is free).  One C-ABI call, asynchronous, for every knob
        combination (estimator mean-pool / sampled-max, block 128 / 64).
        `qk_norm` (optional): (kind, q_weight [H*D], k_weight [H*D], eps) -- the processor's RMSNorm over all heads'
        channels (MW:99-102), applied to q and k inside the gather kernel (include/blade_asa.h: BladeQkNorm).
        `sample_offsets` (sampled_max): (q_off, k_off) int32 [B,H,num_keep]; drawn like W:49-51 when omitted.
        `peers` (optional, _lib.BladePeers): Ulysses pull/push over NVLink peer memory -- q/k/v describe the layout
        only, rows are read from the peers' buffers and the output is written into the peers' buffers (returns
        out = None; the caller owns the barriers)."""
        self._require_cuda(q, k, v, scores)
        B, H, S, D = q.shape
        kn = self.knobs
        dev = q.device
        nb = -(-S // kn.block_size)
        cfg = kn.c_config(nb)
        keep = []                                                   # tensors the C structs point into
        if kn.estimator == "sampled_max" and scores is None:
            assert kn.num_keep == 32, "the estimator kernel is built for num_keep = 32 (W:62)"
            qo, ko = sample_offsets if sample_offsets is not None else (None, None)
            if qo is None:
                qo, ko = self.draw_offsets(B, H, dev), self.draw_offsets(B, H, dev)
            qo, ko = qo.to(torch.int32).contiguous(), ko.to(torch.int32).contiguous()
            assert tuple(qo.shape) == (B, H, kn.num_keep) and tuple(ko.shape) == (B, H, kn.num_keep)
            keep += [qo, ko]
            cfg.sample_q_off, cfg.sample_k_off = qo.data_ptr(), ko.data_ptr()
        sel = selected_acc if selected_acc is not None else self.selected_counter(dev)
        assert sel.dtype == torch.int64 and sel.device == dev
        cfg.selected_acc = sel.data_ptr()
        if qk_norm is not None:
            kind, wq, wk, eps = qk_norm[:4]
            rstd = qk_norm[4] if len(qk_norm) > 4 else None      # statistic computed elsewhere, fp32 [2, B*S] by token
            bq, bk = (qk_norm[5], qk_norm[6]) if len(qk_norm) > 6 else (None, None)   # kind 3: LayerNorm biases
            per_head = int(kind) == 3
            if (virtual_rows is not None or peers is not None) and rstd is None and not per_head:
                raise ValueError("fused q/k norm on the sharded Ulysses layout needs the statistic (qk_rms_stat)")
            n_w = D if per_head else H * D
            assert wq.dtype == q.dtype and wk.dtype == q.dtype and wq.numel() == n_w and wk.numel() == n_w
            wq, wk = wq.contiguous(), wk.contiguous()
            if rstd is not None:
                assert rstd.dtype == torch.float32 and rstd.is_contiguous() and rstd.numel() == 2 * B * S
            if bq is not None:
                assert bq.dtype == q.dtype and bk.dtype == q.dtype and bq.numel() == D and bk.numel() == D
                bq, bk = bq.contiguous(), bk.contiguous()
            norm = BladeQkNorm(int(kind), float(eps), wq.data_ptr(), wk.data_ptr(), ptr(rstd), ptr(bq), ptr(bk))
            keep += [norm, wq, wk, bq, bk, rstd]
            cfg.qk_norm = C.pointer(norm)
        if rope is not None:
            table, first = rope
            assert table.dtype == torch.float32 and table.is_contiguous() and table.shape[-2:] == (D // 2, 2)
            cfg.rope_cos_sin = table.data_ptr()
            cfg.rope_first_row = int(first)
        src = self.src_row(dev, S)
        dst = src
        if peers is not None:
            assert virtual_rows is None and B == 1
            cfg.peers = C.pointer(peers)
            keep.append(peers)
        if virtual_rows is not None:
            # q/k/v rows live at `virtual_rows[s]` of a packed buffer (Ulysses receive layout): compose the gather
            # table with it; the output permutation is unchanged.
            # cached per table OBJECT (a held reference + its version counter): a bare data_ptr key would hand a
            # stale composition to a new tensor that reuses a freed address
            key = ("v", str(dev), S)
            ent = self._src_row.get(key)
            if ent is None or ent[0] is not virtual_rows or ent[1] != virtual_rows._version:
                comp = (virtual_rows[src.long()] if src is not None else virtual_rows).contiguous()
                ent = (virtual_rows, virtual_rows._version, comp)
                self._src_row[key] = ent
            src = ent[2]
            if rope is not None or qk_norm is not None:
                # the rotary table and the norm statistic are indexed by token, not by packed-buffer row
                tok = dst
                if tok is None:
                    tkey = ("arange", str(dev), S)
                    if tkey not in self._src_row:
                        self._src_row[tkey] = torch.arange(S, dtype=torch.int32, device=dev)
                    tok = self._src_row[tkey]
                cfg.token_row = tok.data_ptr()
        nbytes = self.lib.blade_asa_workspace_bytes(B, H, S, D, C.byref(cfg))
        ws = self.workspace(dev, nbytes)
        if peers is not None and out is None:
            out_t = q                                              # layout descriptor only: rows go to peers.out[*]
        else:
            out_t = out if out is not None else torch.empty(B, S, H, D, dtype=q.dtype, device=dev).transpose(1, 2)
        cnt = torch.empty(B, H, nb, dtype=torch.int32, device=dev)
        sc_out = mask = idx = None
        if return_debug:
            n_idx = -(-nb // 2) if kn.block_size == 64 else nb     # block 64: lists over the 128x128 tensor-core tiles
            sc_out = torch.empty(B, H, nb, nb, dtype=torch.float32, device=dev)
            mask = torch.empty(B, H, nb, nb, dtype=torch.uint8, device=dev)
            idx = torch.empty(B, H, n_idx, n_idx, dtype=torch.int32, device=dev)
        if scores is not None:
            scores = scores.contiguous()
            assert scores.dtype == torch.float32 and tuple(scores.shape) == (B, H, nb, nb)
        check(self.lib.blade_asa_forward(
            C.byref(tensor_desc(q)), C.byref(tensor_desc(k)), C.byref(tensor_desc(v)), ptr(src), ptr(dst),
            C.byref(cfg), ptr(scores), C.byref(tensor_desc(out_t)), ptr(sc_out), ptr(mask), ptr(idx), cnt.data_ptr(),
            ws.data_ptr(), ws.numel(), current_stream()))
        del keep
        res = None if (peers is not None and out is None) else out_t
        if return_debug:
            return res, dict(scores=sc_out, mask=mask.bool(), idx=idx, cnt=cnt)
        return res, cnt
