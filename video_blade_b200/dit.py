"""Minimal random-init Wan2.1-T2V-1.3B-shaped DiT + the 8-step distilled sampler, for the end-to-end clip
benchmark of BASELINE.json config 3 (synthetic latents / prompt embeddings, no checkpoint, no diffusers).

This is benchmark scaffolding around the hot path, not part of it: every token-wise op is stock PyTorch/cuBLAS
(as it is in the reference, where diffusers provides them); self-attention goes through the reference-facing
processor + the B200 ASA engine.  Shapes follow diffusers' WanTransformer3DModel for the 1.3B config: 30 blocks,
dim 1536, 12 heads x 128, FFN 8960, text dim 4096 (512 tokens), 16 latent channels, patch (1,2,2), adaLN
modulation from a sinusoidal timestep embedding, 3-axis RoPE (44/42/42 split of the head dim).

Sampler of record: `generate_new` of the reference trainer (train_wanx_tdm.py:1402-1443) with flow sigmas
(shift 3.0, inference.py:49-50): x0 = x_t - sigma*v, eps = x_t + (1-sigma)*v, t <- t - 1000/K,
x_t <- (1-sigma')*x0 + sigma'*(eta*eps + sqrt(1-eta^2)*N(0,1)).

Sequence parallelism: with a UlyssesGroup every rank keeps S/P tokens for all token-wise ops and the
processor wraps `inner_attention` with the head/sequence all-to-all.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .modify_wan import Attention, RMSNorm, WanAttnProcessor2_0, apply_rotary_emb
from .ulysses import UlyssesGroup


class UlyssesWanAttnProcessor(WanAttnProcessor2_0):
    """WanAttnProcessor2_0 (MW:75-148) with the Ulysses exchange around `inner_attention`.

    Fused path (default): q, k, v leave the projections untouched; only the RMSNorm statistic (one float per token
    and tensor) is computed on the token shard and all-gathered, the raw projections travel in ONE all_to_all, and
    the normalisation (my heads' weight slice), the rotary embedding and the curve-order gather all happen inside the
    gather kernel, reading the packed receive buffer in place."""

    def __init__(self, group: Optional[UlyssesGroup], fuse: bool = True, plane=None):
        super().__init__()
        self.group = group
        self.fuse = fuse
        self.plane = plane          # UlyssesPeerPlane: the exchange as NVLink loads/stores inside the layer's kernels

    def _call_peer_plane(self, attn, hidden_states, rotary_emb):
        """Peer-memory data plane: the projections write into the symmetric q/k/v buffer, the RMSNorm statistic of my
        tokens is stored into every peer's table, and the layer pulls / pushes rows over NVLink (UlyssesPeerPlane)."""
        from . import wanx_blocksparseattn as W
        from .modify_wan import _rms_kind, _rope_table
        pl, g = self.plane, self.group
        eng = W._engine(use_rearrange=bool(attn.inner_attention.use_rearrange))
        Hl, D = pl.Hl, pl.D
        sl = slice(g.rank_in_group * Hl * D, (g.rank_in_group + 1) * Hl * D)           # my heads' weights
        outs = []
        for b in range(hidden_states.shape[0]):
            x = hidden_states[b]
            for j, lin in enumerate((attn.to_q, attn.to_k, attn.to_v)):
                dst = pl.qkv[j].view(pl.Sl, pl.H * D)
                if lin.bias is not None:
                    torch.addmm(lin.bias, x, lin.weight.t(), out=dst)
                else:
                    torch.mm(x, lin.weight.t(), out=dst)
            pl.push_rms_stat(eng, attn.norm_q.eps)
            o, cnt = pl.attention(eng, rope=(_rope_table(self._full_rope(rotary_emb)), 0),
                                  qk_norm=(_rms_kind(attn.norm_q), attn.norm_q.weight.detach()[sl],
                                           attn.norm_k.weight.detach()[sl], float(attn.norm_q.eps), pl.rstd.view(-1)),
                                  selected_acc=attn.inner_attention.counter(x.device))
            attn.inner_attention.count_call(cnt)
            o = o.reshape(1, pl.Sl, pl.H * D)
            # B = 1 (the CFG-split clip): to_out reads the symmetric buffer in place.  B > 1: the next sequence's peers
            # overwrite it after their next barrier, so keep a copy (stream-ordered before my side of that barrier)
            outs.append(o if hidden_states.shape[0] == 1 else o.clone())
        o = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return attn.to_out[1](attn.to_out[0](o.type_as(hidden_states)))

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, rotary_emb=None):
        if self.group is None or self.group.P == 1:
            return super().__call__(attn, hidden_states, encoder_hidden_states, attention_mask, rotary_emb)
        import torch.distributed as dist
        from .modify_wan import _norm_fusable, _rms_kind, _rope_table
        g = self.group
        B = hidden_states.shape[0]
        if self.plane is not None and self.fuse and rotary_emb is not None and hidden_states.is_cuda \
                and _norm_fusable(attn, hidden_states, True) and getattr(attn.inner_attention, "supports_fused_rope", False):
            return self._call_peer_plane(attn, hidden_states, rotary_emb)
        q = attn.to_q(hidden_states).unflatten(2, (attn.heads, -1))                    # [B, S/P, H, D]
        k = attn.to_k(hidden_states).unflatten(2, (attn.heads, -1))
        v = attn.to_v(hidden_states).unflatten(2, (attn.heads, -1))
        fused = self.fuse and rotary_emb is not None and _norm_fusable(attn, q.flatten(2, 3), True) \
            and getattr(attn.inner_attention, "supports_fused_rope", False)
        if not fused:
            q = attn.norm_q(q.flatten(2, 3)).unflatten(2, (attn.heads, -1))
            k = attn.norm_k(k.flatten(2, 3)).unflatten(2, (attn.heads, -1))
            if rotary_emb is not None:                                                 # local tokens' freqs
                q = apply_rotary_emb(q.transpose(1, 2), rotary_emb).transpose(1, 2)
                k = apply_rotary_emb(k.transpose(1, 2), rotary_emb).transpose(1, 2)
        outs = []
        for b in range(B):                                                             # exchange is per sequence
            qb, kb, vb = q[b:b + 1].contiguous(), k[b:b + 1].contiguous(), v[b:b + 1].contiguous()
            kw = {}
            if fused:
                from . import wanx_blocksparseattn as W
                eng = W._engine(use_rearrange=bool(attn.inner_attention.use_rearrange))
                stat = eng.qk_rms_stat(qb.transpose(1, 2), kb.transpose(1, 2), attn.norm_q.eps)   # [2, S/P]
                parts = torch.empty(g.P, 2, stat.shape[1], dtype=stat.dtype, device=stat.device)
                dist.all_gather_into_tensor(parts, stat, group=g.group)
                rstd = parts.permute(1, 0, 2).reshape(2, -1).contiguous()              # [2, S] by token
                Hl = attn.heads // g.P
                D = q.shape[-1]
                sl = slice(g.rank_in_group * Hl * D, (g.rank_in_group + 1) * Hl * D)   # my heads' weights
                kw = dict(rotary=(_rope_table(self._full_rope(rotary_emb)), 0),
                          qk_norm=(_rms_kind(attn.norm_q), attn.norm_q.weight.detach()[sl], attn.norm_k.weight.detach()[sl],
                                   float(attn.norm_q.eps), rstd))
            gq, gk, gv, vrow, _keep = self.group.scatter_heads_fused(qb, kb, vb)       # one all_to_all
            o = attn.inner_attention(gq, gk, gv, virtual_rows=vrow, **kw)              # [1, H/P, S, D]
            outs.append(self.group.gather_heads(o.transpose(1, 2)))                    # [1, S/P, H, D]
        o = torch.cat(outs, 0).flatten(2, 3).type_as(hidden_states)
        return attn.to_out[1](attn.to_out[0](o))

    def _full_rope(self, rotary_local):
        """The fused path rotates inside the gather kernel, which sees ALL tokens: it needs the full-sequence table.
        The model hands the processor its shard; `full_rotary_emb` is attached by WanLikeDiT.forward."""
        full = getattr(self, "full_rotary_emb", None)
        assert full is not None, "sequence-parallel fused rotary embedding needs the full-sequence table"
        return full


class UlyssesCogAttnProcessor:
    """SageAttnCogVideoXAttnProcessor (MC:11-76) on a token shard of the concatenated [text ; video] sequence, with
    the Ulysses exchange around `inner_attention`.  `rope_local` = (cos, sin) [S/P, D] for my tokens, identity rows
    for text tokens (MC:59-64 rotates the video part only)."""

    def __init__(self, group: UlyssesGroup, fuse: bool = True, plane=None):
        self.group = group
        self.fuse = fuse
        self.plane = plane          # UlyssesPeerPlane: pull q/k/v rows, push output rows over NVLink
        self._table = None

    def _rope_table(self, rope_full, device):
        cos, sin = rope_full
        key_ = (cos, sin, cos._version, sin._version)   # held references, not bare addresses
        if self._table is None or self._table[0][0] is not cos or self._table[0][1] is not sin \
                or self._table[0][2:] != key_[2:]:
            self._table = (key_, torch.stack([cos[:, 0::2], sin[:, 0::2]], dim=-1).to(device, torch.float32).contiguous())
        return self._table[1]

    def _call_peer_plane(self, attn, hidden_states, rope_full, text_len):
        """Projections written into the symmetric buffer; per-head LayerNorm, rotary embedding (video rows only) and the
        curve-order gather with the text rows at the tail all happen in the kernel that pulls the rows from the peers."""
        from . import cogvideo_blocksparseattn as Cg
        pl = self.plane
        eng = Cg._engine(use_rearrange=bool(attn.inner_attention.use_rearrange))
        nq, nk = attn.norm_q, attn.norm_k
        outs = []
        for b in range(hidden_states.shape[0]):
            x = hidden_states[b]
            for j, lin in enumerate((attn.to_q, attn.to_k, attn.to_v)):
                dst = pl.qkv[j].view(pl.Sl, pl.H * pl.D)
                if lin.bias is not None:
                    torch.addmm(lin.bias, x, lin.weight.t(), out=dst)
                else:
                    torch.mm(x, lin.weight.t(), out=dst)
            o, cnt = pl.attention(eng, rope=(self._rope_table(rope_full, x.device), int(text_len)),
                                  qk_norm=(3, nq.weight.detach(), nk.weight.detach(), float(nq.eps), None,
                                           None if nq.bias is None else nq.bias.detach(),
                                           None if nk.bias is None else nk.bias.detach()),
                                  selected_acc=attn.inner_attention.counter(x.device))
            attn.inner_attention.count_call(cnt)
            o = o.reshape(1, pl.Sl, pl.H * pl.D)
            outs.append(o if hidden_states.shape[0] == 1 else o.clone())
        o = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return attn.to_out[1](attn.to_out[0](o.type_as(hidden_states)))

    def __call__(self, attn, hidden_states, rope_local, rope_full=None, text_len=0):
        from .modify_cogvideo import _norm_fusable, apply_rotary_emb as cog_rope
        B, Sl, _ = hidden_states.shape
        if self.plane is not None and self.fuse and rope_full is not None and hidden_states.is_cuda \
                and _norm_fusable(attn, hidden_states.view(B, Sl, attn.heads, -1), hidden_states.shape[-1] // attn.heads):
            return self._call_peer_plane(attn, hidden_states, rope_full, text_len)
        q = attn.to_q(hidden_states).view(B, Sl, attn.heads, -1)
        k = attn.to_k(hidden_states).view(B, Sl, attn.heads, -1)
        v = attn.to_v(hidden_states).view(B, Sl, attn.heads, -1)
        kw = {}
        if self.fuse and rope_full is not None and _norm_fusable(attn, q, q.shape[-1]):
            # per-head LayerNorm and the rotary embedding need nothing from other ranks: both run inside the gather
            # kernel on the packed receive buffer (my heads, all tokens); the projections travel untouched
            nq, nk = attn.norm_q, attn.norm_k
            kw = dict(rotary=(self._rope_table(rope_full, q.device), int(text_len)),
                      qk_norm=(3, nq.weight.detach(), nk.weight.detach(), float(nq.eps), None,
                               None if nq.bias is None else nq.bias.detach(), None if nk.bias is None else nk.bias.detach()))
        else:
            q = attn.norm_q(q).to(v.dtype)                                             # LayerNorm per head (MC:54-57)
            k = attn.norm_k(k).to(v.dtype)
            q = cog_rope(q.transpose(1, 2), rope_local).transpose(1, 2)
            k = cog_rope(k.transpose(1, 2), rope_local).transpose(1, 2)
        outs = []
        for b in range(B):                                                             # exchange is per sequence
            gq, gk, gv, vrow, _keep = self.group.scatter_heads_fused(q[b:b + 1].contiguous(), k[b:b + 1].contiguous(),
                                                                     v[b:b + 1].contiguous())
            o = attn.inner_attention(gq, gk, gv, virtual_rows=vrow, **kw)              # [1, H/P, S, D], text first
            outs.append(self.group.gather_heads(o.transpose(1, 2)))                    # [1, S/P, H, D]
        o = torch.cat(outs, 0).flatten(2, 3).type_as(hidden_states)
        return attn.to_out[1](attn.to_out[0](o))


class CrossAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.to_out = nn.Linear(dim, dim)
        self.norm_q, self.norm_k = RMSNorm(dim), RMSNorm(dim)

    def forward(self, x, ctx):
        B, S, C = x.shape
        q = self.norm_q(self.to_q(x)).view(B, S, self.heads, -1).transpose(1, 2)
        k = self.norm_k(self.to_k(ctx)).view(B, ctx.shape[1], self.heads, -1).transpose(1, 2)
        v = self.to_v(ctx).view(B, ctx.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)                                    # 512 text keys: not the hot path
        return self.to_out(o.transpose(1, 2).reshape(B, S, C))


class WanBlock(nn.Module):
    def __init__(self, dim, heads, ffn):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6, elementwise_affine=False)
        self.attn1 = Attention(dim, heads, qk_norm="rms_norm_across_heads")
        self.norm2 = nn.LayerNorm(dim, eps=1e-6, elementwise_affine=True)
        self.attn2 = CrossAttention(dim, heads)
        self.norm3 = nn.LayerNorm(dim, eps=1e-6, elementwise_affine=False)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn), nn.GELU(approximate="tanh"), nn.Linear(ffn, dim))
        self.scale_shift_table = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)

    def forward(self, x, ctx, temb, rotary_emb):
        """adaLN block.  The token-wise glue goes through scaffold_ops (one fused pass each on CUDA, the same torch
        expressions as before elsewhere): norm1/norm3 are LayerNorms without affine, eps 1e-6."""
        from . import scaffold_ops as ops
        sh_msa, sc_msa, g_msa, sh_mlp, sc_mlp, g_mlp = (self.scale_shift_table + temb.float()).chunk(6, dim=1)
        h = ops.ln_modulate(x, sc_msa, sh_msa, self.norm1.eps)
        x = ops.gated_residual(x, self.attn1(h, rotary_emb=rotary_emb), g_msa)
        x = x + self.attn2(self.norm2(x), ctx)
        h = ops.ln_modulate(x, sc_mlp, sh_mlp, self.norm3.eps)
        x = ops.gated_residual(x, self.ffn[2](ops.linear_gelu_tanh(h, self.ffn[0])), g_mlp)
        return x


def rope_freqs(frames, height, width, head_dim, theta=10000.0, device="cpu"):
    """Complex rotary table [1,1,S,head_dim/2] over the (f,h,w) token grid, raster order (w fastest)."""
    dh = dw = 2 * (head_dim // 6)
    dt = head_dim - dh - dw

    def axis(n, d):
        inv = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64, device=device) / d))
        ang = torch.outer(torch.arange(n, dtype=torch.float64, device=device), inv)
        return torch.polar(torch.ones_like(ang), ang)
    ft, fh, fw = axis(frames, dt), axis(height, dh), axis(width, dw)
    f = torch.cat([ft[:, None, None].expand(frames, height, width, -1), fh[None, :, None].expand(frames, height, width, -1),
                   fw[None, None, :].expand(frames, height, width, -1)], dim=-1)
    return f.reshape(1, 1, frames * height * width, head_dim // 2).to(torch.complex64)


class WanLikeDiT(nn.Module):
    def __init__(self, dim=1536, heads=12, ffn=8960, layers=30, text_dim=4096, in_ch=16, patch=(1, 2, 2),
                 freq_dim=256):
        super().__init__()
        self.dim, self.heads, self.in_ch, self.patch, self.freq_dim = dim, heads, in_ch, patch, freq_dim
        pdim = in_ch * patch[0] * patch[1] * patch[2]
        self.patch_embedding = nn.Linear(pdim, dim)            # == Conv3d(kernel = stride = patch)
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"), nn.Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, 6 * dim))
        self.blocks = nn.ModuleList([WanBlock(dim, heads, ffn) for _ in range(layers)])
        self.norm_out = nn.LayerNorm(dim, eps=1e-6, elementwise_affine=False)
        self.proj_out = nn.Linear(dim, pdim)
        self.head_table = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)
        self.group: Optional[UlyssesGroup] = None
        self._rope = {}
        self.hoist = False
        self._order = {}

    def set_hoisted_permutation(self, on: bool = True):
        """Model-level hoist of the Gilbert permutation: every token-wise op is permutation equivariant, so
        the tokens (and their rotary table) are put into curve order ONCE per forward, the shared ASA module runs with
        use_rearrange = False (no per-layer gather of q, k, v, no inverse permutation of the output), and the sequence
        is put back before unpatchify.  The reference permutes inside every layer (W:142-159); the drop-in default
        keeps that.  With sequence parallelism the ranks then own contiguous curve segments."""
        self.hoist = bool(on)
        for blk in self.blocks:
            blk.attn1.inner_attention.use_rearrange = not self.hoist

    def _curve_order(self, S, device):
        key = (S, str(device))
        if key not in self._order:
            from . import wanx_blocksparseattn as W
            from .asa import token_order
            order = torch.from_numpy(token_order(W._knobs())).long().to(device)
            assert order.numel() == S, (order.numel(), S)
            inv = torch.empty_like(order)
            inv[order] = torch.arange(S, device=device)
            self._order[key] = (order, inv)
        return self._order[key]

    def set_sequence_parallel(self, group: Optional[UlyssesGroup], data_plane: str = "auto"):
        """data_plane: "p2p" = NVLink peer-memory pull/push inside the layer's kernels (UlyssesPeerPlane, needs
        torch symmetric memory between the group's GPUs), "nccl" = all_to_all around the layer, "auto" = p2p if the
        peer mapping can be set up on every rank, else nccl."""
        self.group = group
        self._plane_mode = data_plane
        self._plane = None
        self.data_plane_in_use = "none" if group is None or group.P == 1 else "nccl"
        for blk in self.blocks:
            blk.attn1.set_processor(UlyssesWanAttnProcessor(group))

    def _ensure_plane(self, Sl, device, dtype):
        """Lazily (the token count is known at the first forward) build ONE peer plane shared by all blocks."""
        g = self.group
        if g is None or g.P == 1 or self._plane_mode == "nccl" or self._plane is not None:
            return
        import torch.distributed as dist
        from .ulysses import UlyssesPeerPlane
        ok, plane = 1, None
        try:
            plane = UlyssesPeerPlane(g, Sl, self.heads, self.dim // self.heads, dtype=dtype, device=device)
        except Exception as e:                                         # no peer mapping: fall back to NCCL everywhere
            ok = 0
            if self._plane_mode == "p2p":
                raise
            print(f"[video_blade_b200] peer-memory plane unavailable ({type(e).__name__}: {e}); using NCCL all_to_all")
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self._plane = plane if int(flag.item()) else False
        if self._plane:
            self.data_plane_in_use = "p2p"
            for blk in self.blocks:
                blk.attn1.get_processor().plane = self._plane

    def patchify(self, lat):                                   # [B,C,F,H,W] -> [B,S,C*p]
        B, C, Fr, H, W = lat.shape
        pt, ph, pw = self.patch
        x = lat.view(B, C, Fr // pt, pt, H // ph, ph, W // pw, pw).permute(0, 2, 4, 6, 1, 3, 5, 7)
        return x.reshape(B, (Fr // pt) * (H // ph) * (W // pw), C * pt * ph * pw), (Fr // pt, H // ph, W // pw)

    def unpatchify(self, x, grid, C):
        B = x.shape[0]
        f, h, w = grid
        pt, ph, pw = self.patch
        x = x.view(B, f, h, w, C, pt, ph, pw).permute(0, 4, 1, 5, 2, 6, 3, 7)
        return x.reshape(B, C, f * pt, h * ph, w * pw)

    def timestep_embedding(self, t):
        half = self.freq_dim // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
        args = t.float()[:, None] * freqs[None]
        return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)

    def forward(self, hidden_states, timestep, encoder_hidden_states):
        """hidden_states [B,16,F,H,W] latents; timestep [B]; encoder_hidden_states [B,512,4096] -> velocity."""
        x, grid = self.patchify(hidden_states)
        S = x.shape[1]
        key = (grid, str(x.device))
        if key not in self._rope:
            self._rope[key] = rope_freqs(*grid, self.dim // self.heads, device=x.device)
        rope = self._rope[key]
        g = self.group
        if self.hoist:                                         # curve order once per forward, not once per layer
            order, inv_order = self._curve_order(S, x.device)
            x, rope = x[:, order], rope[:, :, order]
        if g is not None and g.P > 1:                          # keep my S/P token shard
            sl = slice(g.rank_in_group * (S // g.P), (g.rank_in_group + 1) * (S // g.P))
            self._ensure_plane(S // g.P, x.device, next(self.parameters()).dtype)
            for blk in self.blocks:                            # the fused gather rotates all tokens of my heads
                blk.attn1.get_processor().full_rotary_emb = rope
            x, rope = x[:, sl], rope[:, :, sl]
        x = self.patch_embedding(x)
        temb = self.time_embedding(self.timestep_embedding(timestep).type_as(x))
        tproj = self.time_projection(temb).unflatten(1, (6, -1))
        ctx = self.text_embedding(encoder_hidden_states)
        for blk in self.blocks:
            x = blk(x, ctx, tproj, rope)
        shift, scale = (self.head_table + temb.float().unsqueeze(1)).chunk(2, dim=1)
        x = (self.norm_out(x.float()) * (1 + scale) + shift).type_as(x)
        x = self.proj_out(x)
        if g is not None and g.P > 1:                          # reassemble the sequence for the sampler update
            import torch.distributed as dist
            parts = [torch.empty_like(x) for _ in range(g.P)]
            dist.all_gather(parts, x.contiguous(), group=g.group)
            x = torch.cat(parts, dim=1)
        if self.hoist:
            x = x[:, inv_order]
        return self.unpatchify(x, grid, self.in_ch)


def flow_sigma(t, shift=3.0, total=1000):
    s = t.float() / total
    return shift * s / (1 + (shift - 1) * s)


def make_velocity_fn(transformer, prompt_embeds, negative_embeds=None, guidance_scale=1.0, cfg_ranks=None):
    """Guided velocity v(x_t, T).  Single process: CFG as a batch of two.  `cfg_ranks=(my_branch, cond_rank,
    uncond_rank)`: the two CFG branches live on different rank groups (no traffic until the combine, which is one
    small all_gather of the velocity)."""
    use_cfg = negative_embeds is not None and guidance_scale != 1.0

    def fn(x_t, T):
        if not use_cfg:
            return transformer(x_t, T, prompt_embeds)
        if cfg_ranks is None:
            v2 = transformer(torch.cat([x_t, x_t]), torch.cat([T, T]), torch.cat([prompt_embeds, negative_embeds]))
            v_c, v_u = v2.chunk(2)
        else:
            import torch.distributed as dist
            branch, cond_rank, uncond_rank = cfg_ranks
            v = transformer(x_t, T, prompt_embeds if branch == 0 else negative_embeds).contiguous()
            parts = [torch.empty_like(v) for _ in range(dist.get_world_size())]
            dist.all_gather(parts, v)
            v_c, v_u = parts[cond_rank], parts[uncond_rank]
        return v_u + guidance_scale * (v_c - v_u)
    return fn


@torch.no_grad()
def generate_new(velocity_fn, noise, steps=8, eta=1.0, flow_shift=3.0, total_steps=1000, generator=None):
    """K-step distilled sampling, train_wanx_tdm.py:1402-1443, on a guided-velocity callable."""
    B = noise.shape[0]
    T = torch.full((B,), total_steps - 1, device=noise.device, dtype=torch.long)
    x_t = noise
    latent = noise
    for _ in range(steps):
        v = velocity_fn(x_t, T)
        sigma = flow_sigma(T, flow_shift, total_steps).view(B, 1, 1, 1, 1).to(x_t.dtype)
        latent = x_t - sigma * v                                  # predicted x0            (TW:1426)
        pred_eps = x_t + (1 - sigma) * v                          #                         (TW:1431)
        T = T - total_steps // steps                              #                         (TW:1435)
        add_eps = eta * pred_eps
        if eta != 1.0:
            add_eps = add_eps + ((1 - eta ** 2) ** 0.5) * torch.randn(pred_eps.shape, device=pred_eps.device,
                                                                       dtype=pred_eps.dtype, generator=generator)
        s2 = flow_sigma(T.clamp(min=0), flow_shift, total_steps).view(B, 1, 1, 1, 1).to(x_t.dtype)
        x_t = (1 - s2) * latent + s2 * add_eps                    # add_noise(latent, add_eps, T)  (TW:1437)
    return latent


# ================================================================================================
# CogVideoX-5B-shaped scaffold (BASELINE config 5): 42 blocks, dim 3072, 48 heads x 64, joint text+video
# self-attention (226 text tokens in front, MC:35), adaLN-zero modulation of both streams, 3-axis RoPE on the video
# tokens only.  Same caveats as WanLikeDiT: random init, stock torch ops everywhere except `inner_attention`.
# ================================================================================================
class _LayerNormZero(nn.Module):
    """CogVideoXLayerNormZero: one linear on the time embedding -> shift/scale/gate for video and text."""

    def __init__(self, dim, temb_dim):
        super().__init__()
        self.linear = nn.Linear(temb_dim, 6 * dim)
        self.norm = nn.LayerNorm(dim, eps=1e-5, elementwise_affine=True)

    def forward(self, x, txt, temb):
        from . import scaffold_ops as ops
        sh, sc, g, tsh, tsc, tg = self.linear(F.silu(temb)).chunk(6, dim=1)
        x = ops.ln_modulate(x, sc, sh, self.norm.eps, self.norm.weight, self.norm.bias)      # one fused pass on CUDA
        txt = ops.ln_modulate(txt, tsc, tsh, self.norm.eps, self.norm.weight, self.norm.bias)
        return x, txt, g[:, None], tg[:, None]


class CogBlock(nn.Module):
    def __init__(self, dim, heads, temb_dim):
        super().__init__()
        self.norm1 = _LayerNormZero(dim, temb_dim)
        self.attn1 = Attention(dim, heads, qk_norm="layer_norm")
        self.norm2 = _LayerNormZero(dim, temb_dim)
        self.ff = nn.Sequential(nn.Linear(dim, 4 * dim), nn.GELU(approximate="tanh"), nn.Linear(4 * dim, dim))

    def forward(self, x, txt, temb, rope):
        T = txt.shape[1]
        nx, nt, g, tg = self.norm1(x, txt, temb)
        from . import scaffold_ops as ops
        ax, at = self.attn1(nx, encoder_hidden_states=nt, image_rotary_emb=rope)
        x, txt = ops.gated_residual(x, ax, g), ops.gated_residual(txt, at, tg)
        nx, nt, g, tg = self.norm2(x, txt, temb)
        ff = self.ff[2](ops.linear_gelu_tanh(torch.cat([nt, nx], dim=1), self.ff[0]))
        return ops.gated_residual(x, ff[:, T:], g), ops.gated_residual(txt, ff[:, :T], tg)


def rope_cos_sin(frames, height, width, head_dim, theta=10000.0, device="cpu"):
    """(cos, sin) [S, head_dim] with repeat-interleaved pairs, the layout diffusers hands the CogVideoX processor."""
    f = rope_freqs(frames, height, width, head_dim, theta, device)[0, 0]          # complex [S, D/2]
    return f.real.float().repeat_interleave(2, -1).contiguous(), f.imag.float().repeat_interleave(2, -1).contiguous()


class CogLikeDiT(nn.Module):
    def __init__(self, dim=3072, heads=48, layers=42, text_dim=4096, in_ch=16, patch=2, temb_dim=512):
        super().__init__()
        self.dim, self.heads, self.in_ch, self.patch = dim, heads, in_ch, patch
        self.patch_embedding = nn.Linear(in_ch * patch * patch, dim)
        self.text_proj = nn.Linear(text_dim, dim)
        self.time_embedding = nn.Sequential(nn.Linear(dim, temb_dim), nn.SiLU(), nn.Linear(temb_dim, temb_dim))
        self.transformer_blocks = nn.ModuleList([CogBlock(dim, heads, temb_dim) for _ in range(layers)])
        self.norm_final = nn.LayerNorm(dim, eps=1e-5)
        self.norm_out_linear = nn.Linear(temb_dim, 2 * dim)
        self.norm_out = nn.LayerNorm(dim, eps=1e-5, elementwise_affine=False)
        self.proj_out = nn.Linear(dim, in_ch * patch * patch)
        self._rope = {}
        self.group: Optional[UlyssesGroup] = None

    def set_sequence_parallel(self, group: Optional[UlyssesGroup], data_plane: str = "auto"):
        """Ulysses degree P: every rank keeps S/P tokens of the concatenated [text ; video] sequence.
        data_plane as in WanLikeDiT.set_sequence_parallel."""
        self.group = group
        self._plane_mode = data_plane
        self._plane = None
        self.data_plane_in_use = "none" if group is None or group.P == 1 else "nccl"

    def _ensure_plane(self, Sl, device, dtype):
        g = self.group
        if g is None or g.P == 1 or self._plane_mode == "nccl" or self._plane is not None:
            return
        import torch.distributed as dist
        from .ulysses import UlyssesPeerPlane
        ok, plane = 1, None
        try:
            plane = UlyssesPeerPlane(g, Sl, self.heads, self.dim // self.heads, dtype=dtype, device=device)
        except Exception as e:
            ok = 0
            if self._plane_mode == "p2p":
                raise
            print(f"[video_blade_b200] peer-memory plane unavailable ({type(e).__name__}: {e}); using NCCL all_to_all")
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self._plane = plane if int(flag.item()) else False
        if self._plane:
            self.data_plane_in_use = "p2p"

    def _blocks_sequence_parallel(self, x, txt, temb, rope):
        """The transformer blocks on my shard of [text ; video]; returns the full-length video / text streams."""
        import torch.distributed as dist
        g = self.group
        T, S = txt.shape[1], txt.shape[1] + x.shape[1]
        assert S % g.P == 0, f"{S} tokens do not split over Ulysses degree {g.P}"
        Sl = S // g.P
        sl = slice(g.rank_in_group * Sl, (g.rank_in_group + 1) * Sl)
        hs = torch.cat([txt, x], dim=1)[:, sl]
        is_text = (torch.arange(S, device=x.device)[sl] < T)[None, :, None]
        cos, sin = rope                                                            # [Sv, D] video rows
        D = cos.shape[-1]
        cos_f = torch.cat([torch.ones(T, D, device=cos.device), cos])[sl]         # identity for text rows
        sin_f = torch.cat([torch.zeros(T, D, device=sin.device), sin])[sl]
        self._ensure_plane(Sl, x.device, x.dtype)
        proc = UlyssesCogAttnProcessor(g, plane=self._plane or None)
        from . import scaffold_ops as ops
        n_t = int(is_text.sum())                                # text rows of my shard lead it (text comes first)
        fused = hs.is_cuda and hs.shape[0] == 1                 # B = 1: the two row ranges are contiguous views

        def per_segment(op, *per_token, text, video):
            """op on the text rows with the text parameters and on the video rows with the video parameters."""
            parts = []
            if n_t:
                parts.append(op(*[a[:, :n_t] for a in per_token], *text))
            if n_t < hs.shape[1]:
                parts.append(op(*[a[:, n_t:] for a in per_token], *video))
            return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        for blk in self.transformer_blocks:
            for norm, fn in ((blk.norm1, lambda h, a=blk.attn1: proc(a, h, (cos_f, sin_f), rope_full=rope, text_len=T)),
                             (blk.norm2, lambda h, ff=blk.ff: ff[2](ops.linear_gelu_tanh(h, ff[0])))):
                sh, sc, gt, tsh, tsc, tg = norm.linear(F.silu(temb)).chunk(6, dim=1)
                if fused:
                    ln = norm.norm
                    h = per_segment(lambda x_, s_, b_: ops.ln_modulate(x_, s_, b_, ln.eps, ln.weight, ln.bias), hs,
                                    text=(tsc, tsh), video=(sc, sh))
                    hs = per_segment(lambda x_, y_, g_: ops.gated_residual(x_, y_, g_), hs, fn(h), text=(tg,), video=(gt,))
                    continue
                shift = torch.where(is_text, tsh[:, None], sh[:, None])
                scale = torch.where(is_text, tsc[:, None], sc[:, None])
                gate = torch.where(is_text, tg[:, None], gt[:, None])
                hs = hs + gate * fn(norm.norm(hs) * (1 + scale) + shift)
        parts = [torch.empty_like(hs) for _ in range(g.P)]
        dist.all_gather(parts, hs.contiguous(), group=g.group)
        full = torch.cat(parts, dim=1)
        return full[:, T:], full[:, :T]

    def forward(self, hidden_states, timestep, encoder_hidden_states):
        """hidden_states [B,F,C,H,W] latents (CogVideoX layout); encoder_hidden_states [B,226,4096]."""
        B, Fr, C, H, W = hidden_states.shape
        p = self.patch
        x = hidden_states.view(B, Fr, C, H // p, p, W // p, p).permute(0, 1, 3, 5, 2, 4, 6)
        x = self.patch_embedding(x.reshape(B, Fr * (H // p) * (W // p), C * p * p))
        txt = self.text_proj(encoder_hidden_states)
        half = self.dim // 2
        fr = torch.exp(-math.log(10000.0) * torch.arange(half, device=x.device, dtype=torch.float32) / half)
        args = timestep.float()[:, None] * fr[None]
        temb = self.time_embedding(torch.cat([torch.cos(args), torch.sin(args)], -1).type_as(x))
        key = (Fr, H // p, W // p, str(x.device))
        if key not in self._rope:
            self._rope[key] = rope_cos_sin(Fr, H // p, W // p, self.dim // self.heads, device=x.device)
        rope = self._rope[key]
        if self.group is not None and self.group.P > 1:
            x, txt = self._blocks_sequence_parallel(x, txt, temb, rope)
        else:
            for blk in self.transformer_blocks:
                x, txt = blk(x, txt, temb, rope)
        x = self.norm_final(x)
        sh, sc = self.norm_out_linear(F.silu(temb)).chunk(2, dim=1)
        x = self.proj_out(self.norm_out(x) * (1 + sc[:, None]) + sh[:, None])
        x = x.view(B, Fr, H // p, W // p, C, p, p).permute(0, 1, 4, 2, 5, 3, 6)
        return x.reshape(B, Fr, C, H, W)
