"""Drop-in mirror of the reference's wanx/train/modify_wan.py (MW): the attention processor and installer for
diffusers' WanTransformer3DModel.  Same names, same call signature, same `attn.inner_attention(q,k,v)`
contract (MW:75-168); the inner attention is the B200 ASA engine.

`Attention` below is a duck-typed stand-in for diffusers' attention block (to_q/to_k/to_v, norm_q/norm_k, heads,
to_out, add_k_proj, set_processor/get_processor) so the path runs where diffusers is not installed; with
diffusers present the real `block.attn1` objects are used unchanged.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .wanx_blocksparseattn import AdaptiveBlockSparseAttnTrain


def apply_rotary_emb(hidden_states: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """MW:110-113: complex multiply of interleaved pairs.  The reference upcasts to float64; fp32 is within
    bf16 rounding of it (tests/test_processors.py) and keeps the op off the fp64 pipe."""
    dt = torch.float64 if freqs.dtype == torch.complex128 and not hidden_states.is_cuda else torch.float32
    x = torch.view_as_complex(hidden_states.to(dt).unflatten(3, (-1, 2)))
    f = freqs.to(torch.complex64 if dt == torch.float32 else torch.complex128)
    return torch.view_as_real(x * f).flatten(3, 4).type_as(hidden_states)


_ROPE_TABLES = []     # (freqs tensor held alive, its version counter, fp32 table)


def _rope_table(freqs: torch.Tensor) -> torch.Tensor:
    """complex rotary table [1,1,S,D/2] -> fp32 [S, D/2, 2] of (cos, sin), cached per tensor OBJECT: the entry keeps
    the source tensor alive and checks its version counter, so a freed-and-reused address or an in-place update can
    never return a stale table (a bare data_ptr key could)."""
    for ref, ver, tab in _ROPE_TABLES:
        if ref is freqs and ver == freqs._version:
            return tab
    f = freqs.reshape(freqs.shape[-2], freqs.shape[-1])
    t = torch.stack([f.real, f.imag], dim=-1).to(torch.float32).contiguous()
    if len(_ROPE_TABLES) >= 8:
        _ROPE_TABLES.pop(0)
    _ROPE_TABLES.append((freqs, freqs._version, t))
    return t


class WanAttnProcessor2_0:
    """MW:75-148."""

    def __init__(self, fuse_rope: bool = True, fuse_norm: bool = True):
        self.fuse_rope = fuse_rope
        self.fuse_norm = fuse_norm      # q/k RMSNorm inside the gather kernel instead of ~14 element-wise launches

    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, rotary_emb: Optional[torch.Tensor] = None) -> torch.Tensor:
        encoder_hidden_states_img = None
        if getattr(attn, "add_k_proj", None) is not None:                     # I2V branch, MW:88-91
            encoder_hidden_states_img = encoder_hidden_states[:, :257]
            encoder_hidden_states = encoder_hidden_states[:, 257:]
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states

        query = attn.to_q(hidden_states)                                       # MW:95-97
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        fused_norm = None
        # the norm precedes the rotary embedding (MW:99-116): it may only move into the kernel if the rotation does too
        rope_in_kernel = rotary_emb is None or (self.fuse_rope and getattr(attn.inner_attention, "supports_fused_rope", False)
                                                and getattr(attn, "add_k_proj", None) is None and hidden_states.is_cuda)
        if self.fuse_norm and rope_in_kernel and _norm_fusable(attn, query, encoder_hidden_states is hidden_states):
            nq, nk = attn.norm_q, attn.norm_k
            fused_norm = (_rms_kind(nq), nq.weight.detach(), nk.weight.detach(), float(nq.eps))
        else:
            if attn.norm_q is not None:                                        # MW:99-102
                query = attn.norm_q(query)
            if attn.norm_k is not None:
                key = attn.norm_k(key)
        query = query.unflatten(2, (attn.heads, -1)).transpose(1, 2)           # MW:104-106: strided views
        key = key.unflatten(2, (attn.heads, -1)).transpose(1, 2)
        value = value.unflatten(2, (attn.heads, -1)).transpose(1, 2)
        fused_rope = None
        if rotary_emb is not None:                                             # MW:108-116
            if getattr(attn.inner_attention, "supports_fused_rope", False) and getattr(attn, "add_k_proj", None) is None \
                    and query.is_cuda and self.fuse_rope:
                fused_rope = (_rope_table(rotary_emb), 0)                      # rotated inside the gather kernel
            else:
                query = apply_rotary_emb(query, rotary_emb)
                key = apply_rotary_emb(key, rotary_emb)

        hidden_states_img = None
        if encoder_hidden_states_img is not None:                              # MW:118-131
            key_img = attn.norm_added_k(attn.add_k_proj(encoder_hidden_states_img))
            value_img = attn.add_v_proj(encoder_hidden_states_img)
            key_img = key_img.unflatten(2, (attn.heads, -1)).transpose(1, 2)
            value_img = value_img.unflatten(2, (attn.heads, -1)).transpose(1, 2)
            hidden_states_img = attn.inner_attention(query, key_img, value_img)
            hidden_states_img = hidden_states_img.transpose(1, 2).flatten(2, 3).type_as(query)

        if fused_rope is not None or fused_norm is not None:
            kw = {}
            if fused_rope is not None:
                kw["rotary"] = fused_rope
            if fused_norm is not None:
                kw["qk_norm"] = fused_norm
            hidden_states = attn.inner_attention(query, key, value, **kw)
        else:
            hidden_states = attn.inner_attention(query, key, value)            # MW:135
        hidden_states = hidden_states.transpose(1, 2).flatten(2, 3)            # a view: output memory is [B,S,H,D]
        hidden_states = hidden_states.type_as(query)
        if hidden_states_img is not None:
            hidden_states = hidden_states + hidden_states_img
        hidden_states = attn.to_out[0](hidden_states)                          # MW:146-147
        hidden_states = attn.to_out[1](hidden_states)
        return hidden_states


def _norm_fusable(attn, query, self_attention: bool) -> bool:
    """The fused path covers the Wan2.1 configuration: RMSNorm over all heads' channels on q and k, weights in the
    activation dtype, self-attention, no image branch."""
    nq, nk = getattr(attn, "norm_q", None), getattr(attn, "norm_k", None)
    if nq is None or nk is None or not self_attention or getattr(attn, "add_k_proj", None) is not None:
        return False
    if not getattr(attn.inner_attention, "supports_fused_qk_norm", False) or not query.is_cuda:
        return False
    for n in (nq, nk):
        w = getattr(n, "weight", None)
        if _rms_kind(n) == 0 or w is None or w.dtype != query.dtype or w.numel() != query.shape[-1] \
                or getattr(n, "bias", None) is not None or getattr(n, "eps", None) is None:
            return False
    return _rms_kind(nq) == _rms_kind(nk)


def _rms_kind(n) -> int:
    """Which rounding the module's forward has (BladeQkNorm.kind): 1 = this package's RMSNorm (fp32 chain, one
    rounding), 2 = diffusers.models.normalization.RMSNorm with half-precision weights ((x * rstd) -> weight dtype,
    then * weight), 0 = anything else (e.g. torch.nn.RMSNorm, whose arithmetic is different again): not fused."""
    if isinstance(n, RMSNorm):
        return 2 if getattr(n, "_two_roundings", False) else 1
    if type(n).__name__ == "RMSNorm" and type(n).__module__.startswith("diffusers."):
        return 2
    return 0


def set_adaptive_block_sparse_attn_wanx(model, verbose=False):
    """MW:150-168: one shared ASA module for every block's self-attention."""
    inner_attn = AdaptiveBlockSparseAttnTrain()
    for idx, block in enumerate(model.blocks):
        block.attn1.verbose = verbose
        block.attn1.inner_attention = inner_attn
        origin_processor = block.attn1.get_processor()
        processor = WanAttnProcessor2_0()
        block.attn1.set_processor(processor)
        if not hasattr(block.attn1, "origin_processor"):
            block.attn1.origin_processor = origin_processor
    return inner_attn


# ------------------------------------------------------------------------------------------------
class RMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x):
        from .scaffold_ops import rmsnorm     # fp32 chain, one rounding: one fused pass on CUDA, torch ops elsewhere
        return rmsnorm(x, self.weight, self.eps)


class Attention(nn.Module):
    """Minimal stand-in for diffusers.models.attention_processor.Attention (self-attention use only)."""

    def __init__(self, dim: int, heads: int, qk_norm: str = "rms_norm_across_heads", bias: bool = True,
                 elementwise_affine: bool = True):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=bias)
        self.to_k = nn.Linear(dim, dim, bias=bias)
        self.to_v = nn.Linear(dim, dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim, bias=bias), nn.Dropout(0.0)])
        if qk_norm == "rms_norm_across_heads":          # Wan: RMSNorm over the full inner dim before the head split
            self.norm_q, self.norm_k = RMSNorm(dim), RMSNorm(dim)
        elif qk_norm == "layer_norm":                   # CogVideoX: LayerNorm per head
            hd = dim // heads
            self.norm_q = nn.LayerNorm(hd, eps=1e-6, elementwise_affine=elementwise_affine)
            self.norm_k = nn.LayerNorm(hd, eps=1e-6, elementwise_affine=elementwise_affine)
        else:
            self.norm_q = self.norm_k = None
        self.add_k_proj = None
        self.is_cross_attention = False
        self.inner_attention = None
        self._processor = None

    def set_processor(self, processor):
        self._processor = processor

    def get_processor(self):
        return self._processor

    def forward(self, hidden_states, **kw):
        return self._processor(self, hidden_states, **kw)
