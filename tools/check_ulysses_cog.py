"""torchrun --nproc-per-node P tools/check_ulysses_cog.py : the CogVideoX-shaped scaffold with the token sequence
[text ; video] sharded over P ranks (Ulysses exchange around every ASA call) must reproduce the single-GPU forward
(same weights, same inputs) up to bf16 GEMM re-tiling noise."""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.dit import CogLikeDiT
from video_blade_b200.modify_cogvideo import set_block_sparse_attn_cogvideox
from video_blade_b200.ulysses import UlyssesGroup

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
with torch.device(dev):
    model = CogLikeDiT(layers=2).to(torch.bfloat16).eval()
set_block_sparse_attn_cogvideox(model).print_every = 0
g = torch.Generator().manual_seed(1)
lat = torch.randn(1, 13, 16, 60, 90, generator=g).to(dev, torch.bfloat16)
txt = torch.randn(1, 226, 4096, generator=g).to(dev, torch.bfloat16)
t = torch.full((1,), 500.0, device=dev)
with torch.no_grad():
    ref = model(lat, t, txt)
    ug = UlyssesGroup(world, rank, world)
    for plane in ("nccl", "p2p"):
        model.set_sequence_parallel(ug, data_plane=plane)
        for rep in range(2):
            got = model(lat, t, txt)
        d = (got.float() - ref.float())
        rel = float(d.norm() / ref.float().norm())
        print(f"rank {rank}: ulysses{world} [{model.data_plane_in_use}] vs single GPU rel-L2 {rel:.3e} max|d| "
              f"{float(d.abs().max()):.3e} finite {bool(torch.isfinite(got).all())}", flush=True)
        assert rel < 3e-2 and model.data_plane_in_use == plane
dist.destroy_process_group()
