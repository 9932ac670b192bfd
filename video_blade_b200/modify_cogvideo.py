"""Drop-in mirror of the reference's cogvideox/train/modify_cogvideo.py (MC): processor + installer for
diffusers' CogVideoXTransformer3DModel (MC:11-91).  Text tokens are concatenated in front of the video tokens
(MC:35), q/k are LayerNorm-ed per head (MC:54-57), RoPE touches the video part only (MC:59-64); the ASA module
moves the text rows to the tail internally (C:141-161)."""
from __future__ import annotations

from typing import Optional

import torch

from .cogvideo_blocksparseattn import AdaptiveBlockSparseAttnTrain, standard_attn  # noqa: F401

sparsity_record = []


def apply_rotary_emb(x: torch.Tensor, freqs_cis) -> torch.Tensor:
    """diffusers.models.embeddings.apply_rotary_emb with use_real=True, use_real_unbind_dim=-1 (the CogVideoX
    call at MC:59-64): pairs (x[2i], x[2i+1]) rotated by (cos, sin) of shape [S, D]."""
    cos, sin = freqs_cis
    cos, sin = cos[None, None].to(x.device), sin[None, None].to(x.device)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rot.float() * sin).to(x.dtype)


class SageAttnCogVideoXAttnProcessor:
    """MC:11-76."""

    def __init__(self, idx, fuse_rope: bool = True, fuse_norm: bool = True):
        self.idx = idx
        self.fuse_rope = fuse_rope
        self.fuse_norm = fuse_norm      # per-head LayerNorm of q/k inside the gather kernel
        self._table = None

    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor,
                 attention_mask: Optional[torch.Tensor] = None, image_rotary_emb=None):
        assert attention_mask is None, "Attention mask is not supported"       # MC:31
        text_seq_length = encoder_hidden_states.size(1)
        hidden_states = torch.cat([encoder_hidden_states, hidden_states], dim=1)   # MC:35
        batch_size = hidden_states.shape[0]

        query = attn.to_q(hidden_states)
        key = attn.to_k(hidden_states)
        value = attn.to_v(hidden_states)
        inner_dim = key.shape[-1]
        head_dim = inner_dim // attn.heads
        query = query.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        key = key.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        value = value.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        fused_norm = None
        # the norm precedes the rotary embedding (MC:54-64): it may only move into the kernel if the rotation does too
        rope_in_kernel = image_rotary_emb is None or (self.fuse_rope and not attn.is_cross_attention)
        if self.fuse_norm and rope_in_kernel and _norm_fusable(attn, query, head_dim):
            nq, nk = attn.norm_q, attn.norm_k
            fused_norm = (3, nq.weight.detach(), nk.weight.detach(), float(nq.eps), None,
                          None if nq.bias is None else nq.bias.detach(), None if nk.bias is None else nk.bias.detach())
        else:
            if attn.norm_q is not None:
                query = attn.norm_q(query).to(dtype=value.dtype)
            if attn.norm_k is not None:
                key = attn.norm_k(key).to(dtype=value.dtype)
        fused_rope = None
        if image_rotary_emb is not None:                                       # MC:59-64
            if getattr(attn.inner_attention, "supports_fused_rope", False) and not attn.is_cross_attention \
                    and query.is_cuda and self.fuse_rope:
                cos, sin = image_rotary_emb                                    # [Sv, D], repeat-interleaved pairs
                key_ = (cos, sin, cos._version, sin._version)   # held references, not bare addresses
                if self._table is None or self._table[0][0] is not cos or self._table[0][1] is not sin \
                        or self._table[0][2:] != key_[2:]:
                    tab = torch.stack([cos[:, 0::2], sin[:, 0::2]], dim=-1).to(query.device, torch.float32).contiguous()
                    self._table = (key_, tab)
                fused_rope = (self._table[1], text_seq_length)                 # video rows only
            else:
                query[:, :, text_seq_length:] = apply_rotary_emb(query[:, :, text_seq_length:], image_rotary_emb)
                if not attn.is_cross_attention:
                    key[:, :, text_seq_length:] = apply_rotary_emb(key[:, :, text_seq_length:], image_rotary_emb)
        if fused_rope is not None or fused_norm is not None:
            kw = {}
            if fused_rope is not None:
                kw["rotary"] = fused_rope
            if fused_norm is not None:
                kw["qk_norm"] = fused_norm
            hidden_states = attn.inner_attention(query, key, value, **kw)
        else:
            hidden_states = attn.inner_attention(query, key, value)            # MC:65 (no .contiguous() needed)
        hidden_states = hidden_states.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim)
        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        encoder_hidden_states, hidden_states = hidden_states.split(
            [text_seq_length, hidden_states.size(1) - text_seq_length], dim=1)
        return hidden_states, encoder_hidden_states


def _norm_fusable(attn, query, head_dim: int) -> bool:
    """Fused path: both norms are LayerNorm over the head dimension with parameters in the activation dtype, on a
    self-attention layer; the rotary embedding then has to be fused too (the norm comes first, MC:54-64)."""
    nq, nk = getattr(attn, "norm_q", None), getattr(attn, "norm_k", None)
    if nq is None or nk is None or getattr(attn, "is_cross_attention", False) or not query.is_cuda:
        return False
    if not getattr(attn.inner_attention, "supports_fused_qk_norm", False) \
            or not getattr(attn.inner_attention, "supports_fused_rope", False):
        return False
    for n in (nq, nk):
        if not isinstance(n, torch.nn.LayerNorm) or tuple(n.normalized_shape) != (head_dim,) or n.weight is None \
                or n.weight.dtype != query.dtype or (n.bias is not None and n.bias.dtype != query.dtype):
            return False
    return True


def set_block_sparse_attn_cogvideox(model, verbose=False):
    """MC:79-91."""
    inner_attn = AdaptiveBlockSparseAttnTrain()
    for idx, block in enumerate(model.transformer_blocks):
        block.attn1.verbose = verbose
        block.attn1.inner_attention = inner_attn
        origin_processor = block.attn1.get_processor()
        processor = SageAttnCogVideoXAttnProcessor(idx)
        block.attn1.set_processor(processor)
        if not hasattr(block.attn1, "origin_processor"):
            block.attn1.origin_processor = origin_processor
    return inner_attn
