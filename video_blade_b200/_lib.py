"""ctypes binding of libblade_asa.so (include/blade_asa.h).  No torch types cross the boundary: the
Python side passes raw device pointers, shapes/strides and the current CUDA stream handle.

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BLADE_ASA_LIB") or os.path.join(HERE, "lib", "libblade_asa.so")

BF16, F16, F32, I32, U8 = 0, 1, 2, 3, 4


class BladeTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("shape", C.c_int64 * 4), ("stride", C.c_int64 * 4),
                ("dtype", C.c_int32), ("_pad", C.c_int32)]


class BladeQkNorm(C.Structure):
    _fields_ = [("kind", C.c_int32), ("eps", C.c_float), ("q_weight", C.c_void_p), ("k_weight", C.c_void_p),
                ("rstd", C.c_void_p), ("q_bias", C.c_void_p), ("k_bias", C.c_void_p)]


MAX_PEERS = 8


class BladePeers(C.Structure):
    _fields_ = [("n_peers", C.c_int32), ("my_peer", C.c_int32), ("rows_per_peer", C.c_int32), ("_pad", C.c_int32),
                ("q", C.c_void_p * MAX_PEERS), ("k", C.c_void_p * MAX_PEERS), ("v", C.c_void_p * MAX_PEERS),
                ("out", C.c_void_p * MAX_PEERS)]


class BladeAsaConfig(C.Structure):
    _fields_ = [("block_size", C.c_int32), ("sample_gap", C.c_int32), ("min_retain", C.c_int32),
                ("max_retain", C.c_int32), ("energy_threshold", C.c_float), ("force_last", C.c_int32),
                ("num_keep", C.c_int32), ("estimator", C.c_int32), ("exact_merge", C.c_int32),
                ("rope_first_row", C.c_int32), ("rope_cos_sin", C.c_void_p), ("qk_norm", C.POINTER(BladeQkNorm)),
                ("token_row", C.c_void_p),
                ("sample_q_off", C.c_void_p), ("sample_k_off", C.c_void_p), ("select_rounding", C.c_int32),
                ("_pad2", C.c_int32), ("selected_acc", C.c_void_p), ("peers", C.POINTER(BladePeers))]


# every symbol include/blade_asa.h declares (tests/test_cabi_symbols.py checks the list against the header)
SYMBOLS = [
    "blade_abi_version", "blade_last_error", "blade_device_check", "blade_gilbert_tables",
    "blade_asa_workspace_bytes", "blade_asa_prep", "blade_asa_scores_meanpool", "blade_asa_select",
    "blade_mask_to_index", "blade_block_sparse_attn_fwd", "blade_asa_attn_fwd", "blade_asa_forward",
    "blade_probe_qk", "blade_probe_pv", "blade_profile_events", "blade_asa_sample_tokens", "blade_asa_scores_sampled",
    "blade_mask64_to_index", "blade_block_sparse_attn64_fwd", "blade_asa_attn64_fwd", "blade_asa_prep_rope",
    "blade_attn_workspace_bytes", "blade_qk_rms_stat", "blade_qk_rms_stat_peers",
    "blade_multilevel_pyramid", "blade_multilevel_mask", "blade_level_mask_to_index", "blade_multilevel_attn_fwd",
    "blade_multilevel_bwd_workspace_bytes", "blade_multilevel_attn_bwd",
    "blade_scaffold_ln_modulate", "blade_scaffold_rmsnorm", "blade_scaffold_gated_residual",
    "blade_debug_attn_schedule",
]

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m video_blade_b200.build` "
            "(video_blade_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    T = C.POINTER(BladeTensor)
    CFG = C.POINTER(BladeAsaConfig)
    lib.blade_abi_version.restype = C.c_int
    lib.blade_last_error.restype = C.c_char_p
    lib.blade_device_check.restype = C.c_int
    lib.blade_gilbert_tables.argtypes = [i32, i32, i32, vp, vp]
    lib.blade_asa_workspace_bytes.argtypes = [i64, i64, i64, i64, CFG]
    lib.blade_asa_workspace_bytes.restype = C.c_size_t
    lib.blade_qk_rms_stat.argtypes = [T, T, f32, vp, vp]
    lib.blade_qk_rms_stat_peers.argtypes = [T, T, f32, C.POINTER(vp), i32, i64, i64, vp]
    lib.blade_multilevel_pyramid.argtypes = [T, T, vp, vp, vp, vp, vp, vp, vp]
    lib.blade_multilevel_mask.argtypes = [vp, i64, i64, i64, i64, vp, i32, vp, vp, vp, vp]
    lib.blade_level_mask_to_index.argtypes = [vp, i64, i64, i64, i64, vp, vp, vp]
    lib.blade_scaffold_ln_modulate.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, f32, i32, vp]
    lib.blade_scaffold_rmsnorm.argtypes = [vp, vp, vp, i64, i64, f32, i32, vp]
    lib.blade_scaffold_gated_residual.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, vp]
    lib.blade_multilevel_bwd_workspace_bytes.argtypes = [i64, i64, i64, i64, i64]
    lib.blade_multilevel_bwd_workspace_bytes.restype = C.c_size_t
    lib.blade_multilevel_attn_bwd.argtypes = [T, T, T, T, T, T, T, T, T, vp, vp, i64, T, T, vp, f32, T, T, T, vp, C.c_size_t, vp]
    lib.blade_multilevel_attn_fwd.argtypes = [T, T, T, T, T, T, T, T, T, vp, vp, i64, T, vp, vp, f32, vp, C.c_size_t, vp]
    lib.blade_attn_workspace_bytes.argtypes = [i64]
    lib.blade_attn_workspace_bytes.restype = C.c_size_t
    lib.blade_asa_prep.argtypes = [T, T, T, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.blade_asa_prep_rope.argtypes = [T, T, T, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, i32, vp]
    lib.blade_asa_scores_meanpool.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp]
    lib.blade_asa_select.argtypes = [vp, i64, i64, i64, i64, CFG, vp, vp, vp, vp, vp, vp, vp]
    lib.blade_mask_to_index.argtypes = [vp, i64, i64, i64, i64, vp, vp, vp]
    lib.blade_block_sparse_attn_fwd.argtypes = [T, T, T, vp, vp, i64, T, vp, vp, f32, vp, C.c_size_t, vp]
    lib.blade_asa_attn_fwd.argtypes = [T, T, T, vp, vp, i64, T, T, i32, T, vp, f32, i32, vp, C.c_size_t, vp]
    lib.blade_asa_forward.argtypes = [T, T, T, vp, vp, CFG, vp, T, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.blade_profile_events.argtypes = [i32, vp, vp]
    lib.blade_probe_qk.argtypes = [vp, vp, vp, i32, vp]
    lib.blade_debug_attn_schedule.argtypes = [i64, i64, i64, i32, i32, vp, i32, i32, vp, i64, vp]
    lib.blade_probe_pv.argtypes = [vp, vp, vp, i32, vp]
    lib.blade_mask64_to_index.argtypes = [vp, i64, i64, i64, i64, vp, vp, vp]
    lib.blade_block_sparse_attn64_fwd.argtypes = lib.blade_block_sparse_attn_fwd.argtypes
    lib.blade_asa_attn64_fwd.argtypes = lib.blade_asa_attn_fwd.argtypes
    lib.blade_asa_sample_tokens.argtypes = [T, T, vp, vp, vp, vp, i32, vp]
    lib.blade_asa_scores_sampled.argtypes = [vp, vp, vp, i64, i64, i64, i64, i32, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name, None)
        if fn is not None and fn.restype is C.c_int and name not in ("blade_abi_version",):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        msg = load().blade_last_error().decode(errors="replace")
        raise RuntimeError(f"blade_asa error {code}: {msg}")


def _dtype_code(t) -> int:
    import torch
    return {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32, torch.int32: I32, torch.uint8: U8}[t.dtype]


class TensorLayout:
    """A [B,H,S,D] layout descriptor without storage behind all of it: base pointer, shape, element strides, dtype.
    Used where the rows of a logical tensor live in several GPUs' memories (Ulysses peer pull / push): the kernels take
    the per-peer base pointers from BladePeers and only the geometry from here."""

    def __init__(self, ptr: int, shape, stride, dtype, device):
        self._ptr, self.shape, self._stride, self.dtype, self.device = int(ptr), tuple(shape), tuple(stride), dtype, device
        self.is_cuda = True

    def dim(self):
        return len(self.shape)

    def stride(self, i=None):
        return self._stride if i is None else self._stride[i]

    def data_ptr(self):
        return self._ptr


def tensor_desc(t) -> BladeTensor:
    """[B,H,S,D] torch tensor (any strides, last dim contiguous) or TensorLayout -> BladeTensor."""
    assert t.dim() == 4, "expected [B,H,S,D]"
    d = BladeTensor()
    d.ptr = t.data_ptr()
    for i in range(4):
        d.shape[i] = t.shape[i]
        d.stride[i] = t.stride(i)
    d.dtype = _dtype_code(t)
    return d


def ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
