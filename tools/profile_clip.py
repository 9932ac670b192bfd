"""torch.profiler breakdown of one Wan-shaped DiT forward (few layers) to see what surrounds the ASA call."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.dit import WanLikeDiT
from video_blade_b200.modify_wan import set_adaptive_block_sparse_attn_wanx
dev = torch.device("cuda", 0)
torch.manual_seed(0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = WanLikeDiT(layers=L).to(dev, torch.bfloat16).eval()
set_adaptive_block_sparse_attn_wanx(model).print_every = 0
lat = torch.randn(2, 16, 21, 60, 104, device=dev, dtype=torch.bfloat16)
txt = torch.randn(2, 512, 4096, device=dev, dtype=torch.bfloat16)
t = torch.full((2,), 500.0, device=dev)
with torch.no_grad():
    for _ in range(2):
        model(lat, t, txt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); model(lat, t, txt); e1.record(); torch.cuda.synchronize()
    print(f"forward (B=2, {L} layers): {e0.elapsed_time(e1):.2f} ms -> {e0.elapsed_time(e1)/L:.2f} ms per layer")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(lat, t, txt)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
