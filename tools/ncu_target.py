"""Short ncu target: a few whole-layer calls at a BASELINE shape with on-device random inputs."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.asa import AsaEngine, AsaKnobs
model = sys.argv[1] if len(sys.argv) > 1 else "wan"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kn = AsaKnobs.wan() if model == "wan" else AsaKnobs.cog()
if len(sys.argv) > 3:
    kn.estimator = sys.argv[3]          # e.g. sampled_max
H, D = (12, 128) if model == "wan" else (48, 64)
S = kn.width * kn.height * kn.depth + kn.text_length
torch.manual_seed(0)
q, k, v = (torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16).transpose(1, 2) for _ in range(3))
eng = AsaEngine(kn)
so = None
if kn.estimator == "sampled_max":
    so = (eng.draw_offsets(1, H, q.device), eng.draw_offsets(1, H, q.device))
for _ in range(n):
    out, cnt = eng.forward(q, k, v, sample_offsets=so)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()), float(cnt.float().mean()))
