"""TEST INFRASTRUCTURE.  Golden vectors for oracle/multilevel.py from the REFERENCE's own code, run in the build
container (needs /root/reference; not needed on the GPU box):

  * the multi-level Triton forward kernel K9 (`_forward`, BLOCK_M = BLOCK_N = POOLING_BLOCK_N = 128 as N:10
    instantiates it) under TRITON_INTERPRET=1 on CPU, fp32 inputs;
  * `transfer_attn_to_mask` of N (multi-level mask from block scores), whose source is executed stand-alone
    (the module itself imports CUDA-only pieces).

    python -m oracle.make_golden_multilevel      ->  tests/golden/multilevel.npz
"""
import ast
import importlib.util
import os
import sys

os.environ["TRITON_INTERPRET"] = "1"        # before triton is imported

import numpy as np
import torch

REF = "/root/reference/cogvideox/sample_evaluate/Triton"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_k9():
    spec = importlib.util.spec_from_file_location(
        "k9_ref", os.path.join(REF, "kernels", "block_sparse_attn_kernel_with_backward_9_10.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.is_hip = lambda: False

    class _NoDevice:                          # `with torch.cuda.device(q.device.index)` on CPU tensors
        def __init__(self, *a):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False
    torch.cuda.device = _NoDevice
    return mod


def load_mask_fn():
    """transfer_attn_to_mask of N, compiled from its own source text (N:154-207), sort pinned to stable."""
    src = open(os.path.join(REF, "cogvideo_newattn.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "transfer_attn_to_mask")
    code = compile(ast.Module(body=[fn], type_ignores=[]), "cogvideo_newattn.py", "exec")

    class _Torch:                             # torch with a stable sort (the reference leaves tie order open)
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def sort(x, dim=-1, descending=False):
            return torch.sort(x, dim=dim, descending=descending, stable=True)
    ns = {"torch": _Torch()}
    exec(code, ns)
    # the module-level `mask_ratios` the reference actually runs with (N:13-19; differs from the function's default)
    assign = next(n for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "mask_ratios")
    return ns["transfer_attn_to_mask"], ast.literal_eval(assign.value)


def golden_backward(k9, out, g):
    """K9's backward kernels (K9:695-1237, 1375-1576) under the interpreter: dq, dk, dv for a mask with every level;
    the gradient reaches k and v through the pooled copies as well."""
    B, H, N, D = 1, 1, 384, 64
    q, k, v = (torch.randn(B, H, N, D, generator=g).requires_grad_(True) for _ in range(3))
    mask = torch.tensor([[1, 2, 0], [4, 1, 8], [2, 8, 1]], dtype=torch.int32).view(1, 1, 3, 3)
    fn = k9.sparse_attention_factory(BLOCK_M=128, BLOCK_N=128)
    o = fn(q, k, v, mask, None)
    do = torch.randn(B, H, N, D, generator=g)
    o.backward(do)
    for key, t in (("q", q), ("k", k), ("v", v), ("mask", mask), ("do", do), ("o", o), ("dq", q.grad), ("dk", k.grad),
                   ("dv", v.grad)):
        out[f"bwd_{key}"] = t.detach().numpy()
    print("backward ok", float(q.grad.abs().mean()), float(k.grad.abs().mean()), float(v.grad.abs().mean()))


def golden_layer(out):
    """The reference module N end to end on CPU: its own GilbertRearranger, random token sampling (draws recorded),
    Triton estimator and Triton multi-level kernel under the interpreter, fp32."""
    import importlib
    import types
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    sys.path.insert(0, os.path.dirname(REF))
    try:
        N = importlib.import_module("Triton.cogvideo_newattn")
    finally:
        sys.path.remove(os.path.dirname(REF))
    for m in list(sys.modules.values()):
        if getattr(m, "__name__", "").startswith("Triton.kernels") and hasattr(m, "is_hip"):
            m.is_hip = lambda: False
    g = torch.Generator().manual_seed(7)
    draws = []

    class _Torch:                              # N's `torch`: no device='cuda', stable sort, recorded random draws
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def tensor(*a, **k):
            k.pop("device", None)
            return torch.tensor(*a, **k)

        @staticmethod
        def sort(*a, **k):
            k["stable"] = True
            return torch.sort(*a, **k)

        @staticmethod
        def rand(*shape, **k):
            k.pop("device", None)
            r = torch.rand(*shape, generator=g)
            draws.append(r.clone())
            return r
    N.torch = _Torch()
    grid, text = (8, 8, 8), 64
    N.width, N.height, N.depth, N.text_length = grid[0], grid[1], grid[2], text
    mod = N.AdaptiveBlockSparseAttnTrain()
    B, H, D = 1, 2, 64
    S = grid[0] * grid[1] * grid[2] + text
    q, k, v = (torch.randn(B, H, S, D, generator=g) for _ in range(3))
    o = mod(q, k, v)
    assert len(draws) == 2                     # one draw for q, one for k (N:84-85), shared by all blocks
    for key, t in (("q", q), ("k", k), ("v", v), ("o", o), ("rand_q", draws[0]), ("rand_k", draws[1])):
        out[f"layer_{key}"] = t.numpy()
    out["layer_grid_text"] = np.array(list(grid) + [text], dtype=np.int64)
    out["layer_ratios"] = np.array([[lv, a, b] for lv, (a, b) in N.mask_ratios.items()], dtype=np.float64)
    print("layer ok", tuple(o.shape), float(o.abs().mean()))


def main():
    k9 = load_k9()
    mask_fn, module_ratios = load_mask_fn()
    out = {}
    g = torch.Generator().manual_seed(42)
    # ---- masks
    for name, (B, H, n) in {"m_small": (1, 2, 16), "m_cog": (1, 2, 139)}.items():
        attn = torch.softmax(torch.randn(B, H, n, n, generator=g) * 2, dim=-1)
        attn[0, 0, 3, :4] = attn[0, 0, 3, 0]                      # ties
        out[f"{name}_attn"] = attn.numpy()
        out[f"{name}_mask"] = mask_fn(attn.clone()).numpy()
        out[f"{name}_mask_module_ratios"] = mask_fn(attn.clone(), module_ratios).numpy()
    out["module_ratios"] = np.array([[lv, a, b] for lv, (a, b) in module_ratios.items()], dtype=np.float64)
    # ---- attention
    cases = {"a_levels": (1, 2, 512, 64), "a_ragged": (1, 1, 300, 64), "a_d128": (1, 1, 384, 128)}
    for name, (B, H, N, D) in cases.items():
        q, k, v = (torch.randn(B, H, N, D, generator=g) for _ in range(3))
        nb = -(-N // 128)
        mask = torch.randint(0, 5, (B, H, nb, nb), generator=g)
        mask = torch.tensor([0, 1, 2, 4, 8], dtype=torch.int32)[mask]
        mask[..., torch.arange(nb), torch.arange(nb)] = 1         # every row attends something
        o = k9._forward(None, q, k, v, mask.contiguous(), 1.0 / D ** 0.5, BLOCK_M=128, BLOCK_N=128, POOLING_BLOCK_N=128)
        for key, t in (("q", q), ("k", k), ("v", v), ("mask", mask), ("o", o)):
            out[f"{name}_{key}"] = t.numpy()
        print(name, "ok", tuple(o.shape), float(o.abs().mean()))
    golden_backward(k9, out, g)
    golden_layer(out)
    np.savez_compressed(os.path.join(OUT, "multilevel.npz"), **out)
    print("wrote", os.path.join(OUT, "multilevel.npz"))


if __name__ == "__main__":
    main()
