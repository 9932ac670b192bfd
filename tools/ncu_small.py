"""ncu target for the small mask kernels (diagnostic)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.asa import AsaEngine, AsaKnobs
kn = AsaKnobs.wan(); eng = AsaEngine(kn)
torch.manual_seed(0)
qm = torch.randn(1, 12, 256, 128, device="cuda"); km = torch.randn(1, 12, 256, 128, device="cuda")
for _ in range(3):
    sc = eng.scores_meanpool(qm, km)
    idx, cnt, _ = eng.select(sc, want_mask=False)
torch.cuda.synchronize(); print("ok", float(cnt.float().mean()))
