"""video_blade_b200 -- B200-native Adaptive Sparse Attention (ASA) hot path of Video-BLADE.

Hand-written sm_100a CUDA kernels behind a C ABI (include/blade_asa.h, video_blade_b200/csrc/), and a
Python host layer that mirrors the reference's attention-processor API:

    from video_blade_b200.modify_wan import set_adaptive_block_sparse_attn_wanx
    from video_blade_b200.modify_cogvideo import set_block_sparse_attn_cogvideox

There is no CPU path and no PyTorch fallback: every compute call goes through libblade_asa.so.
"""
from . import _lib  # noqa: F401
from .asa import AsaKnobs, AsaEngine  # noqa: F401

__all__ = ["AsaKnobs", "AsaEngine"]
