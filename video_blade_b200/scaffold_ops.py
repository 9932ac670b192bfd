"""Fused token-wise glue for the benchmark scaffolds (dit.py) -- NOT part of the ASA hot path.

The random-init DiTs that the 8-step clip benchmark times spend (in eager PyTorch) more time in adaLN modulation,
gated residuals and dtype casts than in attention or in the GEMMs: each `(norm(x.float()) * (1 + scale) + shift)
.type_as(x)` is 5-7 passes over [B,S,C] in fp32.  These three wrappers run each of them as ONE pass
(csrc/scaffold_kernels.cu) on CUDA bf16/fp16 tensors and fall back to the plain torch expression everywhere else (CPU
tests of the scaffold logic), so the scaffold's arithmetic is the same up to where the roundings happen."""
from __future__ import annotations

import torch

from . import _lib


def _fusable(x, C):
    return x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and C % 256 == 0 and C <= 4096


def ln_modulate(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, eps: float, weight=None, bias=None) -> torch.Tensor:
    """LayerNorm over the last dim in fp32 (optional affine weight / bias), then `* (1 + scale) + shift` (scale / shift
    [B,1,C] or [B,C]) -> x.dtype."""
    B, S, C = x.shape
    if not _fusable(x, C) or (weight is not None and weight.dtype != x.dtype):
        h = torch.nn.functional.layer_norm(x.float(), (C,), None if weight is None else weight.float(),
                                           None if bias is None else bias.float(), eps=eps)
        return (h * (1 + scale.reshape(B, 1, C).float()) + shift.reshape(B, 1, C).float()).type_as(x)
    x = x.contiguous()
    sc = scale.reshape(B, C).float().contiguous()
    sh = shift.reshape(B, C).float().contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        w = None if weight is None else weight.contiguous()
        b = None if bias is None else bias.contiguous()
        _lib.check(_lib.load().blade_scaffold_ln_modulate(x.data_ptr(), sc.data_ptr(), sh.data_ptr(), _lib.ptr(w), _lib.ptr(b),
                                                          out.data_ptr(), B, S, C, float(eps), _lib._dtype_code(x),
                                                          _lib.current_stream()))
    return out


def gated_residual(x: torch.Tensor, y: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """x + y * gate in fp32 (gate [B,1,C] or [B,C], fp32) -> x.dtype."""
    B, S, C = x.shape
    if not _fusable(x, C) or y.dtype != x.dtype:
        return (x.float() + y.float() * gate.reshape(B, 1, C).float()).type_as(x)
    x, y = x.contiguous(), y.contiguous()
    g = gate.reshape(B, C).float().contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().blade_scaffold_gated_residual(x.data_ptr(), y.data_ptr(), g.data_ptr(), out.data_ptr(), B, S, C,
                                                             _lib._dtype_code(x), _lib.current_stream()))
    return out


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """x * rsqrt(mean(x^2) + eps) * weight over the last dim, fp32 arithmetic, one rounding."""
    C = x.shape[-1]
    if not _fusable(x, C) or weight.dtype != x.dtype or torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        v = x.float()
        v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)
        return (v * weight.float()).type_as(x)
    xc = x.contiguous()
    out = torch.empty_like(xc)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().blade_scaffold_rmsnorm(xc.data_ptr(), weight.contiguous().data_ptr(), out.data_ptr(),
                                                      xc.numel() // C, C, float(eps), _lib._dtype_code(x), _lib.current_stream()))
    return out


def linear_gelu_tanh(x: torch.Tensor, lin: torch.nn.Linear) -> torch.Tensor:
    """gelu_tanh(x @ W^T + b) with the activation in the GEMM epilogue (cuBLASLt) on CUDA."""
    if x.is_cuda and lin.bias is not None and hasattr(torch, "_addmm_activation"):
        shp = x.shape
        y = torch._addmm_activation(lin.bias, x.reshape(-1, shp[-1]), lin.weight.t(), use_gelu=True)
        return y.view(*shp[:-1], lin.out_features)
    return torch.nn.functional.gelu(lin(x), approximate="tanh")
