"""ORACLE tooling -- generate tests/golden/* by running the REFERENCE's own Python code.

Run in the build container only (needs /root/reference, read-only):
    python -m oracle.make_golden

How the reference is executed without its absent third-party dependencies:
  * `block_sparse_attn`, `block_sparse_attn.bert_padding`, `matplotlib(.pyplot)` are stubbed in
    sys.modules so that W (wanx_blocksparseattn.py) and C (cogvideo_blocksparseattn.py) import;
  * the module-global `block_sparse_attn` (the wrapper around the external CUDA kernel, W:278-309)
    is replaced by the definitional dense-masked function from oracle/asa_oracle.py -- the external
    kernel itself cannot be built here, so that ONE function is "parity unpinned";
  * the Triton estimator `attn_with_pooling` cannot target a CPU; it is replaced by its fp32/bf16
    restatement, and the restatement itself is pinned against the real Triton kernel run under
    TRITON_INTERPRET=1 (fp32, small shape) in `estimator_*.npz`;
  * C's GilbertRearranger hard-codes device='cuda' (C:127-128); the module's `torch` name is
    proxied so `torch.tensor(..., device='cuda')` lands on the CPU;
  * `torch.sort` is called unstably by the reference (W:217).  For tie-free score rows the result is
    order independent and the reference function is run untouched; for the tie-heavy fixtures the
    module's `torch.sort` is forced to stable=True (the canonical order this project pins).
Everything else -- pad_to_multiple, random_sample_tokens, simple_pooling, transfer_attn_to_mask,
GilbertRearranger, the LSE merge, AdaptiveBlockSparseAttnTrain.forward -- is reference code, unmodified.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _stub_modules():
    bsa = types.ModuleType("block_sparse_attn")
    bsa.block_sparse_attn_func = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("external kernel absent"))
    bp = types.ModuleType("block_sparse_attn.bert_padding")
    bp.pad_input = lambda *a, **k: None
    bp.unpad_input = lambda *a, **k: None
    bsa.bert_padding = bp
    sys.modules["block_sparse_attn"] = bsa
    sys.modules["block_sparse_attn.bert_padding"] = bp
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


class _TorchProxy:
    """Stands in for the `torch` name inside a reference module; strips device='cuda'."""

    def __init__(self, stable_sort=False):
        self._stable = stable_sort

    def __getattr__(self, name):
        return getattr(torch, name)

    def tensor(self, *a, **k):
        k.pop("device", None)
        return torch.tensor(*a, **k)

    def sort(self, *a, **k):
        if self._stable:
            k["stable"] = True
        return torch.sort(*a, **k)


def load_reference(tree: str, modname: str):
    """Import special_attentions_local.TrainRelated.<modname> from <REF>/<tree>/train."""
    _stub_modules()
    for m in [m for m in sys.modules if m.startswith("special_attentions_local")]:
        del sys.modules[m]
    path = os.path.join(REF, tree, "train")
    sys.path.insert(0, path)
    try:
        mod = importlib.import_module(f"special_attentions_local.TrainRelated.{modname}")
    finally:
        sys.path.remove(path)
    return mod


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def from_bf16_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)


# ---------------------------------------------------------------------------------------------
def golden_gilbert(W):
    grids = [(52, 30, 21), (45, 30, 13), (52, 30, 5), (8, 6, 4), (2, 2, 2), (3, 5, 7), (26, 15, 4),
             (15, 10, 6), (13, 10, 6), (1, 1, 5), (7, 1, 1), (4, 9, 2), (5, 5, 5), (2, 3, 11)]
    res = {}
    for (w, h, d) in grids:
        rr = W.GilbertRearranger(w, h, d, 0)
        c2r = rr.original_order2gilbert_order.numpy().astype(np.int64)
        r2c = rr.gilbert_order2original_order.numpy().astype(np.int64)
        res[f"{w}x{h}x{d}"] = {
            "sha256_curve2raster": hashlib.sha256(c2r.tobytes()).hexdigest(),
            "sha256_raster2curve": hashlib.sha256(r2c.tobytes()).hexdigest(),
            "head": c2r[:16].tolist(),
        }
    with open(os.path.join(OUT, "gilbert_hashes.json"), "w") as f:
        json.dump(res, f, indent=1)
    print("gilbert:", len(res), "grids")


def _score_rows(B, H, nb, seed, kind):
    g = torch.Generator().manual_seed(seed)
    if kind == "softmax":        # tie-free: peaked softmax rows
        x = torch.randn(B, H, nb, nb, generator=g) * 2.5
        return torch.softmax(x, dim=-1)
    if kind == "flat":           # near-uniform: selection saturates at max_retain
        x = torch.rand(B, H, nb, nb, generator=g) + 4.0
        return x / x.sum(-1, keepdim=True)
    if kind == "peaky":          # a handful of dominant blocks: selection hits min_retain
        x = torch.randn(B, H, nb, nb, generator=g) * 8.0
        return torch.softmax(x, dim=-1)
    if kind == "ties":           # quantised scores -> long exact-tie runs
        x = torch.randint(0, 6, (B, H, nb, nb), generator=g).float() + 1.0
        return x / x.sum(-1, keepdim=True)
    raise ValueError(kind)


def golden_select(W, C):
    arrays = {}
    meta = []
    case = 0
    for nb in (61, 122, 139, 256):
        for kind in ("softmax", "flat", "peaky", "ties"):
            for flavor in ("wan", "cog"):
                B, H = 1, 2
                sc = _score_rows(B, H, nb, 1000 + case, kind)
                stable = kind in ("ties",)
                mod = W if flavor == "wan" else C
                saved_torch = mod.torch
                mod.torch = _TorchProxy(stable_sort=stable)
                try:
                    if flavor == "wan":
                        mx, mn = 0.17, 0.05
                        mask = mod.transfer_attn_to_mask(sc.clone(), mode="energy", init_k=None,
                                                         max_retain_ratio=mx, min_retain_ratio=mn,
                                                         energy_threshold=0.95)
                    else:
                        mx, mn = 0.1, 0.05
                        mxt = torch.ones([B, H]) * mx      # C:347-348
                        mnt = torch.ones([B, H]) * mn
                        mask = mod.transfer_attn_to_mask(sc.clone(), mode="energy", init_k=None,
                                                         max_retain_ratio=mxt, min_retain_ratio=mnt,
                                                         energy_threshold=0.95)
                finally:
                    mod.torch = saved_torch
                arrays[f"scores_{case}"] = sc.numpy()
                arrays[f"mask_{case}"] = np.packbits(mask.numpy(), axis=-1)
                meta.append(dict(case=case, nb=nb, kind=kind, flavor=flavor, max_ratio=mx, min_ratio=mn,
                                 thr=0.95, stable_sort_forced=stable))
                case += 1
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "select_cases.npz"), **arrays)
    print("select:", case, "cases")


def golden_helpers(W):
    """pad_to_multiple / random_sample_tokens / simple_pooling / efficient path pieces."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 2, 300, 16, generator=g).bfloat16()
    arrays = {"x": bf16_bits(x)}
    arrays["pad128"] = bf16_bits(W.pad_to_multiple(x, 128))
    arrays["pool30"] = bf16_bits(W.simple_pooling(x, sample_gap=30))
    arrays["pool15"] = bf16_bits(W.simple_pooling(x, sample_gap=15))
    xp = W.pad_to_multiple(x, 128)
    torch.manual_seed(11)
    arrays["sampled"] = bf16_bits(W.random_sample_tokens(xp, 128, 32))
    np.savez_compressed(os.path.join(OUT, "helpers.npz"), **arrays)
    print("helpers: ok")


def golden_estimator(W=None):
    """Real Triton kernel (P) under TRITON_INTERPRET=1 (must be set before triton is imported, so this
    runs in a child process), fp32 inputs, small shape."""
    import subprocess
    env = dict(os.environ, TRITON_INTERPRET="1")
    r = subprocess.run([sys.executable, "-m", "oracle.make_golden", "--estimator-child"], env=env,
                       cwd=os.path.dirname(OUT.rstrip("/")).rsplit("/tests", 1)[0],
                       capture_output=True, text=True)
    print(r.stdout.strip()[-400:] or r.stderr.strip()[-400:])


def _estimator_child():
    W = load_reference("wanx", "wanx_blocksparseattn")
    P = sys.modules["special_attentions_local.TrainRelated.attn_pooling_kernel"]
    P.is_hip = lambda: False
    g = torch.Generator().manual_seed(5)
    nb, nk, D = 6, 32, 64
    sq = torch.randn(1, 2, nb * nk, D, generator=g)
    sk = torch.randn(1, 2, nb * nk, D, generator=g)
    v = torch.zeros_like(sq)
    _, po = P.attn_with_pooling(sq, sk, v, False, 1.0 / (D ** 0.5), nk)
    np.savez_compressed(os.path.join(OUT, "estimator_triton_fp32.npz"),
                        sq=sq.numpy(), sk=sk.numpy(), po=po.numpy())
    print("estimator (triton interpreter): ok", tuple(po.shape))


def _run_reference_layer(mod, flavor, q, k, v, grid, text_length, gap, max_ratio, min_ratio, seed):
    from oracle import asa_oracle as O

    def dense_sub(q_, k_, v_, block_mask):
        out, lse = O.dense_masked_attention(q_, k_, v_, block_mask, block_q=128, block_k=128)
        return out, lse.unsqueeze(-1).to(q_.dtype)               # shape/dtype of W:309

    def est_sub(sq, sk, v_, causal, sm_scale, num_keep):
        # restatement of P on already-sampled tokens (block == num_keep, offsets = identity)
        B, H, Ls, D = sq.shape
        ident = torch.arange(num_keep).view(1, 1, num_keep).expand(B, H, num_keep)
        po = O.estimator_sampled_max(sq, sk, num_keep, ident, ident)
        return None, po

    mod.block_sparse_attn = dense_sub
    mod.attn_with_pooling = est_sub
    mod.width, mod.height, mod.depth = grid
    mod.text_length = text_length
    mod.sample_gap = gap
    mod.max_retain_ratio = max_ratio
    mod.min_retain_ratio = min_ratio
    # simple_pooling's default argument was bound at import; adaptive_block_sparse_attn passes
    # sample_gap explicitly (W:344-345) so the module global is what counts.
    saved_torch = mod.torch
    mod.torch = _TorchProxy(stable_sort=True)
    try:
        layer = mod.AdaptiveBlockSparseAttnTrain()
        torch.manual_seed(seed)                                   # feeds torch.rand in W:50
        out = layer(q, k, v)
        # second call with use_rearrange disabled
        layer.use_rearrange = False
        torch.manual_seed(seed)
        out_nr = layer(q, k, v)
    finally:
        mod.torch = saved_torch
    return out, out_nr


def golden_layers(W, C):
    from oracle import asa_oracle as O
    # --- wan flavour: 26x15x4 grid = 1560 tokens (ragged: 12 full blocks + 24), D=128
    grid = (26, 15, 4)
    S = grid[0] * grid[1] * grid[2]
    q, k, v = O.synth_qkv(1, 2, S, 128, seed=21, structured=2.0, grid=grid)
    out, out_nr = _run_reference_layer(W, "wan", q, k, v, grid, 0, 30, 0.4, 0.05, seed=3)
    np.savez_compressed(os.path.join(OUT, "layer_wan_small.npz"),
                        q=bf16_bits(q), k=bf16_bits(k), v=bf16_bits(v),
                        out=bf16_bits(out), out_norearrange=bf16_bits(out_nr),
                        meta=np.frombuffer(json.dumps(dict(
                            grid=grid, text_length=0, sample_gap=30, max_retain_ratio=0.4,
                            min_retain_ratio=0.05, rng_seed=3, estimator="sampled_max")).encode(), np.uint8))
    print("layer wan:", tuple(out.shape), float(out.float().abs().mean()))
    # --- cog flavour: 15x10x6 = 900 video tokens + 40 text tokens first, D=64
    grid = (15, 10, 6)
    T = 40
    S = grid[0] * grid[1] * grid[2] + T
    q, k, v = O.synth_qkv(1, 3, S, 64, seed=22, structured=2.0, grid=grid, text_length=T)
    out, out_nr = _run_reference_layer(C, "cog", q, k, v, grid, T, 15, 0.3, 0.05, seed=4)
    np.savez_compressed(os.path.join(OUT, "layer_cog_small.npz"),
                        q=bf16_bits(q), k=bf16_bits(k), v=bf16_bits(v),
                        out=bf16_bits(out), out_norearrange=bf16_bits(out_nr),
                        meta=np.frombuffer(json.dumps(dict(
                            grid=grid, text_length=T, sample_gap=15, max_retain_ratio=0.3,
                            min_retain_ratio=0.05, rng_seed=4, estimator="sampled_max")).encode(), np.uint8))
    print("layer cog:", tuple(out.shape), float(out.float().abs().mean()))


def main():
    if not os.path.isdir(REF):
        raise SystemExit("make_golden needs /root/reference (build container only)")
    os.makedirs(OUT, exist_ok=True)
    W = load_reference("wanx", "wanx_blocksparseattn")
    golden_gilbert(W)
    golden_helpers(W)
    golden_estimator(W)
    W = load_reference("wanx", "wanx_blocksparseattn")
    Wmods = {m: sys.modules[m] for m in list(sys.modules) if m.startswith("special_attentions_local")}
    C = load_reference("cogvideox", "cogvideo_blocksparseattn")
    golden_select(W, C)
    golden_layers(W, C)
    del Wmods


if __name__ == "__main__":
    if "--estimator-child" in sys.argv:
        _estimator_child()
    else:
        main()
