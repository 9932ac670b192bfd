"""ORACLE (test infrastructure, not product code) -- Adaptive Sparse Attention, CPU restatement.

Plain torch-on-CPU restatement of Video-BLADE's ASA hot path.  Every function cites the reference
lines it follows.  Abbreviations (all under /root/reference/):
  W  = wanx/train/special_attentions_local/TrainRelated/wanx_blocksparseattn.py
  C  = cogvideox/train/special_attentions_local/TrainRelated/cogvideo_blocksparseattn.py
  P  = wanx/train/special_attentions_local/TrainRelated/attn_pooling_kernel.py

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product (video_blade_b200) never does.

Pinning status
  * Everything except `block_sparse_attn_func` is pinned against the reference's own Python code,
    imported in the build container by oracle/make_golden.py (fixtures in tests/golden/).
  * `block_sparse_attn_func` (W:301-305) lives in the third-party CUDA library
    mit-han-lab/Block-Sparse-Attention (installed from git HEAD, unpinned: reference README.md:52-61)
    whose source is not in /root/reference and which no reference test exercises.  Its restatement
    here (`dense_masked_attention`) is definitional: softmax(QK^T/sqrt(D)) restricted to the selected
    blocks, times V, plus the fp32 log-sum-exp.  **parity unpinned** for that one function.

Determinism rules the reference leaves open and this oracle fixes:
  * estimator sample offsets are explicit inputs (the reference draws torch.rand per call, W:50);
  * block selection sorts with value descending, block index ascending (torch.sort stable=True);
  * prefix sums follow torch's CPU cumsum: fp64 sequential accumulation, each prefix rounded to fp32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .gilbert import gilbert_permutations

LOG2E = 1.44269504  # the literal the Triton kernel uses (P:163)


# --------------------------------------------------------------------------------------------
# configuration (module-level knobs of the reference, W:9-16 / C:9-16, gathered in one record)
# --------------------------------------------------------------------------------------------
@dataclass
class ASAConfig:
    flavor: str = "wan"            # "wan" | "cog"
    use_rearrange: bool = True     # W:9
    max_retain_ratio: float = 0.17  # W:10 (0.1 for cog, C:10)
    min_retain_ratio: float = 0.05  # W:11
    width: int = 52                # W:12
    height: int = 30               # W:13
    depth: int = 21                # W:14
    sample_gap: int = 30           # W:15 (15 for cog)
    text_length: int = 0           # W:16 (226 for cog)
    block_size: int = 128          # literal W:325
    num_keep: int = 32             # literal W:62
    energy_threshold: float = 0.95  # literal W:341
    estimator: str = "meanpool"    # "meanpool" (north-star kernel (a)) | "sampled_max" (reference P)

    @staticmethod
    def wan(**kw):
        return ASAConfig(flavor="wan", **kw)

    @staticmethod
    def cog(**kw):
        base = dict(flavor="cog", max_retain_ratio=0.1, width=45, height=30, depth=13,
                    sample_gap=15, text_length=226)
        base.update(kw)
        return ASAConfig(**base)


# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------
def pad_to_multiple(x: torch.Tensor, multiple: int) -> torch.Tensor:
    """W:25-36 -- pad the sequence dim (2) to a multiple by replicating the last token."""
    L = x.size(2)
    r = L % multiple
    if r == 0:
        return x
    last = x[:, :, -1:, :].expand(-1, -1, multiple - r, -1)
    return torch.cat([x, last], dim=2)


def retain_bounds(nb: int, min_ratio: float, max_ratio: float, flavor: str = "wan") -> Tuple[int, int]:
    """min/max retained blocks per row.

    wan: `max(1, int(seq * ratio))` in Python double arithmetic (W:215-216).
    cog: `clamp((seq * ratio_tensor).to(int), min=1)` where ratio_tensor is an fp32 tensor
         (`ones * ratio`, C:347-348) -- the product is an fp32 multiply, truncated (C:230-231).
    """
    if flavor == "cog":
        lo = int(np.float32(nb) * np.float32(min_ratio))
        hi = int(np.float32(nb) * np.float32(max_ratio))
    else:
        lo = int(nb * min_ratio)
        hi = int(nb * max_ratio)
    return max(1, lo), max(1, hi)


def draw_sample_offsets(B: int, H: int, block_size: int, num_keep: int,
                        generator: torch.Generator) -> torch.Tensor:
    """W:49-51 -- one set of `num_keep` intra-block offsets per (b, h), shared by every block:
    rand[B,H,1,block] -> topk indices.  Returned int64 [B,H,num_keep] in topk (descending-rand) order."""
    rand_vals = torch.rand(B, H, 1, block_size, generator=generator)
    _, idx = torch.topk(rand_vals, num_keep, dim=3)
    return idx[:, :, 0, :]


def sample_tokens(x: torch.Tensor, block_size: int, offsets: torch.Tensor) -> torch.Tensor:
    """W:37-60 with the random offsets made explicit.  x [B,H,L,D], L % block == 0."""
    B, H, L, D = x.shape
    nb = L // block_size
    xb = x.reshape(B, H, nb, block_size, D)
    idx = offsets[:, :, None, :, None].expand(B, H, nb, offsets.size(-1), D)
    return torch.gather(xb, 3, idx).reshape(B, H, nb * offsets.size(-1), D)


# --------------------------------------------------------------------------------------------
# block-score estimators
# --------------------------------------------------------------------------------------------
def estimator_sampled_max(q, k, block_size, q_offsets, k_offsets, out_dtype=None):
    """The reference estimator: W:62-87 -> P:201-253 (Triton `_attn_fwd`, non-causal STAGE=3 path).

    With sampled q~,k~ (num_keep tokens per block) and s = (q~ . k~) * (1/sqrt(D)) * LOG2E:
      R[r, j]  = max_{c in k-block j} s[r, c]           stored in q.dtype      (P:54-55)
      m[r]     = max_j (fp32 R before the cast)                                 (P:52-56)
      Po[i, j] = max_{r in q-block i} exp2(R[r, j] - m[r])  stored in q.dtype   (P:72-82; l_i == 1)
      Po      /= Po.sum(-1)                                 in q.dtype          (P:250-251)
    Returns Po [B,H,nb,nb] in `out_dtype` (default q.dtype).
    """
    out_dtype = out_dtype or q.dtype
    qp = pad_to_multiple(q, block_size)
    kp = pad_to_multiple(k, block_size)
    sq = sample_tokens(qp, block_size, q_offsets)
    sk = sample_tokens(kp, block_size, k_offsets)
    B, H, Ls, D = sq.shape
    nk = q_offsets.size(-1)
    nb = Ls // nk
    scale = (1.0 / (D ** 0.5)) * LOG2E
    s = torch.matmul(sq.float(), sk.float().transpose(-1, -2))              # fp32 accumulate
    bm = s.reshape(B, H, Ls, nb, nk).amax(-1) * scale                        # [B,H,Ls,nb] fp32
    m = bm.amax(-1, keepdim=True)
    R = bm.to(out_dtype).float()
    po = torch.exp2(R - m).reshape(B, H, nb, nk, nb).amax(3).to(out_dtype)   # [B,H,nb,nb]
    ssum = po.sum(-1, keepdim=True)
    return po / ssum


def estimator_meanpool(q, k, block_size):
    """North-star kernel (a) (BASELINE.json): block-mean-pool Q and K (over the replicate-padded
    sequence, W:25-36), coarse block-score GEMM scaled by 1/sqrt(D), row softmax.  fp32 throughout.
    Returns fp32 [B,H,nb,nb].  (No reference counterpart: the reference estimator is the sampled-max
    one above; SURVEY.md section 7.3 records why both exist.)"""
    qp = pad_to_multiple(q, block_size).float()
    kp = pad_to_multiple(k, block_size).float()
    B, H, L, D = qp.shape
    nb = L // block_size
    qm = qp.reshape(B, H, nb, block_size, D).mean(3)
    km = kp.reshape(B, H, nb, block_size, D).mean(3)
    s = torch.matmul(qm, km.transpose(-1, -2)) * (1.0 / (D ** 0.5))
    return torch.softmax(s, dim=-1)


# --------------------------------------------------------------------------------------------
# block selection
# --------------------------------------------------------------------------------------------
def select_blocks_energy(scores: torch.Tensor, min_retain, max_retain,
                         energy_threshold: float = 0.95, force_last: int = 0):
    """`transfer_attn_to_mask(mode="energy")`, W:214-229 (wan) / C:228-248 (cog).

    scores [B,H,nq,nk] (any float dtype; fp32 in the parity contract).
    min_retain / max_retain: ints, or int tensors [B,H] (cog's per-head bounds, C:230-231).
    force_last: cog ORs the last two block rows and columns to True (C:247-248) -> force_last=2.

    Returns (mask bool [B,H,nq,nk], k_sel int64 [B,H,nq]) where k_sel is the clamped cut.
    Canonical tie order: value descending, index ascending (stable sort).
    """
    B, H, nq, nk = scores.shape
    sorted_attn, indices = torch.sort(scores, dim=-1, descending=True, stable=True)
    cum = torch.cumsum(sorted_attn, dim=-1)
    total = cum[..., -1:]
    thr = energy_threshold * total
    energy_mask = cum >= thr
    k_idx = torch.argmax(energy_mask.int(), dim=-1)
    unsatisfied = (cum[..., -1:] < thr).squeeze(-1)
    k_idx = torch.where(unsatisfied, torch.full_like(k_idx, nk), k_idx)
    if torch.is_tensor(min_retain):
        lo = min_retain.to(k_idx.dtype).reshape(B, H, 1)
        hi = max_retain.to(k_idx.dtype).reshape(B, H, 1)
        k_idx = torch.minimum(torch.maximum(k_idx, lo), hi)
    else:
        k_idx = torch.clamp(k_idx, min=int(min_retain), max=int(max_retain))
    pos = torch.arange(nk).view(1, 1, 1, nk)
    keep = pos < k_idx.unsqueeze(-1)
    mask = torch.zeros(B, H, nq, nk, dtype=torch.bool)
    mask.scatter_(-1, indices, keep)
    if force_last:
        mask[:, :, :, -force_last:] = True
        mask[:, :, -force_last:, :] = True
    return mask, k_idx


def mask_to_index_list(mask: torch.Tensor):
    """Compact per-row block-index list (north-star kernel (a) output): ascending block indices,
    -1 padded to nk entries, plus per-row counts.  idx int32 [B,H,nq,nk], cnt int32 [B,H,nq]."""
    B, H, nq, nk = mask.shape
    cnt = mask.sum(-1).to(torch.int32)
    ar = torch.arange(nk).view(1, 1, 1, nk).expand(B, H, nq, nk)
    key = torch.where(mask, ar, ar + nk)
    order = torch.sort(key, dim=-1).values
    idx = torch.where(order < nk, order, torch.full_like(order, -1)).to(torch.int32)
    return idx, cnt


# --------------------------------------------------------------------------------------------
# attention (the external block_sparse_attn_func, restated definitionally -- parity unpinned)
# --------------------------------------------------------------------------------------------
def dense_masked_attention(q, k, v, block_mask, block_q=128, block_k=None, row_chunk=2048):
    """`block_sparse_attn(q,k,v,block_mask)` W:278-309: out = softmax(QK^T/sqrt(D) + mask)V over the
    True blocks of `block_mask` [B,H,ceil(Sq/bq)(+),ceil(Sk/bk)(+)]; extra mask rows/cols (the
    reference's `S//128+1` quirk, W:22) are cropped.  Literal dense-masked evaluation in fp32.

    Returns (out in q.dtype [B,H,Sq,D], lse fp32 [B,H,Sq])."""
    block_k = block_k or block_q
    B, H, Sq, D = q.shape
    Sk = k.size(2)
    scale = 1.0 / (D ** 0.5)
    out = torch.empty(B, H, Sq, D, dtype=torch.float32)
    lse = torch.empty(B, H, Sq, dtype=torch.float32)
    kcol_blk = torch.arange(Sk) // block_k
    kf = k.float()
    vf = v.float()
    for r0 in range(0, Sq, row_chunk):
        r1 = min(Sq, r0 + row_chunk)
        qrow_blk = torch.arange(r0, r1) // block_q
        tok_mask = block_mask[:, :, qrow_blk][:, :, :, kcol_blk]              # [B,H,r,Sk]
        s = torch.matmul(q[:, :, r0:r1].float(), kf.transpose(-1, -2)) * scale
        s = s.masked_fill(~tok_mask, float("-inf"))
        l = torch.logsumexp(s, dim=-1)
        p = torch.exp(s - l.unsqueeze(-1))
        out[:, :, r0:r1] = torch.matmul(p, vf)
        lse[:, :, r0:r1] = l
    return out.to(q.dtype), lse


def block_gather_attention(q, k, v, idx, cnt, block=128):
    """Same function as `dense_masked_attention`, evaluated only over the selected blocks (identical
    mathematics because masked columns contribute exp(-inf) = 0; different fp32 summation order).
    Fast enough for full-size heads.  idx/cnt from `mask_to_index_list`."""
    B, H, Sq, D = q.shape
    Sk = k.size(2)
    nq = (Sq + block - 1) // block
    scale = 1.0 / (D ** 0.5)
    out = torch.empty(B, H, Sq, D, dtype=torch.float32)
    lse = torch.empty(B, H, Sq, dtype=torch.float32)
    for b in range(B):
        for h in range(H):
            kf = k[b, h].float()
            vf = v[b, h].float()
            for i in range(nq):
                r0, r1 = i * block, min(Sq, (i + 1) * block)
                blks = idx[b, h, i, : int(cnt[b, h, i])].long()
                cols = (blks[:, None] * block + torch.arange(block)[None, :]).reshape(-1)
                cols = cols[cols < Sk]
                s = (q[b, h, r0:r1].float() @ kf[cols].T) * scale
                l = torch.logsumexp(s, dim=-1)
                out[b, h, r0:r1] = torch.exp(s - l[:, None]) @ vf[cols]
                lse[b, h, r0:r1] = l
    return out.to(q.dtype), lse


def simple_pooling(x: torch.Tensor, sample_gap: int) -> torch.Tensor:
    """W:88-93 -- replicate-pad to a multiple of `sample_gap`, mean over each group (result in x.dtype)."""
    x = pad_to_multiple(x, sample_gap)
    B, H, L, D = x.shape
    return x.reshape(B, H, L // sample_gap, sample_gap, D).mean(dim=-2)


def standard_attn(q, k_pool, v_pool):
    """W:21-24 -- dense attention of every query over the pooled keys (all-ones block mask)."""
    B, H = q.shape[:2]
    ones = torch.ones(B, H, 1, 1, dtype=torch.bool)
    return dense_masked_attention(q, k_pool, v_pool, ones, block_q=q.size(2), block_k=k_pool.size(2))


def merge_lse(out1, lse1, out2, lse2, sample_gap: int):
    """W:351-370 -- blend of the sparse branch and the pooled branch, op by op in the tensors' dtype.
    lse1/lse2 arrive as [B,H,S,1] already cast to q.dtype (W:309)."""
    gap_t = torch.tensor(sample_gap, dtype=lse1.dtype)
    log_gap = torch.log(gap_t)
    lw1 = lse1
    lw2 = lse2 + log_gap
    mx = torch.maximum(lw1, lw2)
    e1 = torch.exp(lw1 - mx)
    e2 = torch.exp(lw2 - mx)
    alpha = e1 / (e1 + e2)
    return out1 * alpha + out2 * (1 - alpha)


# --------------------------------------------------------------------------------------------
# Gilbert rearrangement
# --------------------------------------------------------------------------------------------
class GilbertRearranger:
    """W:102-159 (wan) / C:110-161 (cog: text tokens first on input, moved to the tail)."""

    def __init__(self, width, height, depth, text_length=0):
        c2r, r2c = gilbert_permutations(width, height, depth)
        self.curve2raster = torch.from_numpy(c2r)
        self.raster2curve = torch.from_numpy(r2c)
        self.text_length = text_length

    def rearrange(self, x):
        if self.text_length:
            t, vid = x[..., : self.text_length, :], x[..., self.text_length:, :]
            return torch.cat((vid.index_select(-2, self.curve2raster), t), dim=-2)
        return x.index_select(-2, self.curve2raster)

    def reversed_rearrange(self, out):
        if self.text_length:
            vid, t = out[..., : -self.text_length, :], out[..., -self.text_length:, :]
            return torch.cat((t, vid.index_select(-2, self.raster2curve)), dim=-2)
        return out.index_select(-2, self.raster2curve)


# --------------------------------------------------------------------------------------------
# the layer
# --------------------------------------------------------------------------------------------
@dataclass
class ASAResult:
    out: torch.Tensor
    sparsity: float
    scores: torch.Tensor
    mask: torch.Tensor
    out1: Optional[torch.Tensor] = None
    lse1: Optional[torch.Tensor] = None
    out2: Optional[torch.Tensor] = None
    lse2: Optional[torch.Tensor] = None
    extra: dict = field(default_factory=dict)


def block_scores(q, k, cfg: ASAConfig, q_offsets=None, k_offsets=None, scores=None):
    if scores is not None:
        return scores
    if cfg.estimator == "meanpool":
        return estimator_meanpool(q, k, cfg.block_size)
    if cfg.estimator == "sampled_max":
        return estimator_sampled_max(q, k, cfg.block_size, q_offsets, k_offsets)
    raise ValueError(cfg.estimator)


def select_mask(scores, cfg: ASAConfig):
    nb = scores.size(-1)
    lo, hi = retain_bounds(nb, cfg.min_retain_ratio, cfg.max_retain_ratio, cfg.flavor)
    force = 2 if cfg.flavor == "cog" else 0
    return select_blocks_energy(scores, lo, hi, cfg.energy_threshold, force_last=force)


def adaptive_block_sparse_attn(q, k, v, cfg: ASAConfig, q_offsets=None, k_offsets=None,
                               scores=None, fast=False) -> ASAResult:
    """W:311-372 / C:327-394 -- q,k,v already in Gilbert order, [B,H,S,D]."""
    sc = block_scores(q, k, cfg, q_offsets, k_offsets, scores)
    mask, _ = select_mask(sc, cfg)
    if fast:
        idx, cnt = mask_to_index_list(mask)
        out1, lse1 = block_gather_attention(q, k, v, idx, cnt, cfg.block_size)
    else:
        out1, lse1 = dense_masked_attention(q, k, v, mask, cfg.block_size)
    lse1 = lse1.unsqueeze(-1).to(q.dtype)                                   # W:309
    k_pool = simple_pooling(k, cfg.sample_gap)
    v_pool = simple_pooling(v, cfg.sample_gap)
    out2, lse2 = standard_attn(q, k_pool, v_pool)
    lse2 = lse2.unsqueeze(-1).to(q.dtype)
    out = merge_lse(out1, lse1, out2, lse2, cfg.sample_gap)
    sparsity = float(1 - mask.float().mean() - 1 / cfg.sample_gap)          # W:372
    return ASAResult(out=out, sparsity=sparsity, scores=sc, mask=mask,
                     out1=out1, lse1=lse1, out2=out2, lse2=lse2)


def asa_forward(q, k, v, cfg: ASAConfig, q_offsets=None, k_offsets=None, scores=None,
                fast=False, rearranger: Optional[GilbertRearranger] = None) -> ASAResult:
    """`AdaptiveBlockSparseAttnTrain.forward` W:383-408 / C:405-427 (bookkeeping prints omitted)."""
    if cfg.use_rearrange:
        rr = rearranger or GilbertRearranger(cfg.width, cfg.height, cfg.depth, cfg.text_length)
        q_r, k_r, v_r = rr.rearrange(q), rr.rearrange(k), rr.rearrange(v)
    else:
        rr = None
        q_r, k_r, v_r = q, k, v
    res = adaptive_block_sparse_attn(q_r, k_r, v_r, cfg, q_offsets, k_offsets, scores, fast)
    res.extra["out_r"] = res.out
    if rr is not None:
        res.out = rr.reversed_rearrange(res.out)
    return res


# --------------------------------------------------------------------------------------------
# RoPE as the Wan processor applies it (modify_wan.py:108-116): complex multiply in float64
# --------------------------------------------------------------------------------------------
def apply_rotary_emb_wan(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """x [B,H,S,D] real; freqs complex128 [1,1,S,D/2].  Pairs are interleaved (x[2i], x[2i+1])."""
    xr = torch.view_as_complex(x.to(torch.float64).unflatten(3, (-1, 2)))
    return torch.view_as_real(xr * freqs).flatten(3, 4).type_as(x)


# --------------------------------------------------------------------------------------------
# synthetic inputs and FLOP accounting are shared with bench.py: they live in the package (not test code)
# --------------------------------------------------------------------------------------------
from video_blade_b200.synth import attention_flops, synth_qkv  # noqa: E402,F401
