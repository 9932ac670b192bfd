"""Timing of the multi-level pooled sparse attention at the reference's own test shape
(test_block_sparse_attention.py:171-186: B=1, H=4, N=17776, D=64, bf16) and at the CogVideoX layer shape (48 heads), with
the module's ratio table (N:10-19: 5 % full, 10 % 2x, 10 % 4x, 25 % 8x, 50 % skipped) and the default one (N:172-178).
Prints one JSON line per configuration: forward / backward ms (CUDA events, L2 flushed), algorithmic TFLOP/s."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200 import cogvideo_newattn as N

e = N._engine()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


for H in (4, 48):
    for name, ratios in (("module", N.mask_ratios), ("default", None)):
        torch.manual_seed(123)
        B, S, D = 1, 17776, 64
        q, k, v, do = (torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16).transpose(1, 2) for _ in range(4))
        nb = -(-S // 128)
        scores = torch.softmax(torch.randn(B, H, nb, nb, device="cuda") * 2.0, -1)
        mask, idx, cnt4 = e.level_mask(scores, ratios, 2)
        pyr = e.pyramid(k, v)
        out, lse = e.attention(q, k, v, pyr, idx, cnt4, want_lse=True)
        keys = (cnt4.float() * torch.tensor([128.0, 64.0, 32.0, 16.0], device="cuda")).sum(-1)      # keys per q-row block
        flops = float(4.0 * D * 128.0 * keys.sum())                                                  # QK^T + PV
        t_pyr = timed(lambda: e.pyramid(k, v))
        t_mask = timed(lambda: e.level_mask(scores, ratios, 2))
        t_fwd = timed(lambda: e.attention(q, k, v, pyr, idx, cnt4))
        t_bwd = timed(lambda: e.attention_bwd(q, k, v, pyr, idx, cnt4, out, lse, do))
        tiles = float((cnt4[..., 0] + (cnt4[..., 1] + 1) // 2 + (cnt4[..., 2] + 3) // 4 + (cnt4[..., 3] + 7) // 8).float().mean())
        print(json.dumps({"shape": [B, H, S, D], "ratios": name, "tiles_per_row_mean": tiles,
                          "pyramid_ms": t_pyr, "level_mask_ms": t_mask, "forward_ms": t_fwd, "backward_ms": t_bwd,
                          "forward_tflops": flops / t_fwd / 1e9, "backward_tflops": 2.5 * flops / t_bwd / 1e9,
                          "note": "backward = first correct version (one tile in flight, fp32 atomics), 2.5x the forward FLOPs"}))
