"""Small target for compute-sanitizer: one ASA layer (wan + cog flavours, ragged tails) + estimator + block-64."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.asa import AsaEngine, AsaKnobs
torch.manual_seed(0)
for kn, H, D in [(AsaKnobs.wan(width=13, height=10, depth=4, max_retain_ratio=0.4), 2, 128),
                 (AsaKnobs.cog(width=9, height=10, depth=4, text_length=40, max_retain_ratio=0.3), 2, 64),
                 (AsaKnobs.wan(width=13, height=10, depth=4, max_retain_ratio=0.4, estimator="sampled_max"), 1, 128),
                 (AsaKnobs.wan(width=13, height=10, depth=4, max_retain_ratio=0.4, block_size=64), 1, 128)]:
    S = kn.width * kn.height * kn.depth + kn.text_length
    q, k, v = (torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16).transpose(1, 2) for _ in range(3))
    out, cnt = AsaEngine(kn).forward(q, k, v)
    torch.cuda.synchronize()
    print(kn.flavor, kn.estimator, kn.block_size, S, float(out.float().abs().mean()), int(torch.isnan(out.float()).sum()))
print("ok")
