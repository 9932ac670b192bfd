"""torchrun --nproc-per-node P tools/check_ulysses_proc.py : the sequence-parallel Wan processor (statistic
all-gather + raw q/k/v all_to_all + norm / rotary / gather fused in the gather kernel) against the single-GPU
processor with torch RMSNorm and torch rotary embedding on the full sequence."""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200 import wanx_blocksparseattn as W
from video_blade_b200.dit import UlyssesWanAttnProcessor, rope_freqs
from video_blade_b200.modify_wan import Attention, WanAttnProcessor2_0
from video_blade_b200.ulysses import UlyssesGroup

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
W.width, W.height, W.depth = 26, 16, 8
W.max_retain_ratio = W.min_retain_ratio = 1.0   # keep every block: block selection on near-uniform random scores
                                                # flips on rounding noise and would mask what is compared here
S, heads, D = 26 * 16 * 8, 4 * world, 128
dim = heads * D
torch.manual_seed(0)
attn = Attention(dim, heads, qk_norm="rms_norm_across_heads").to(dev, torch.bfloat16)
with torch.no_grad():
    attn.norm_q.weight.copy_(1 + 0.3 * torch.randn(dim))
    attn.norm_k.weight.copy_(1 + 0.3 * torch.randn(dim))
inner = W.AdaptiveBlockSparseAttnTrain(); inner.print_every = 0
attn.inner_attention = inner
x = torch.randn(2, S, dim, device=dev, dtype=torch.bfloat16)
rope = rope_freqs(8, 16, 26, D, device=dev)
with torch.no_grad():
    attn.set_processor(WanAttnProcessor2_0(fuse_rope=False, fuse_norm=False))
    ref = attn(x, rotary_emb=rope).float()
    ug = UlyssesGroup(world, rank, world)
    sl = slice(rank * (S // world), (rank + 1) * (S // world))
    for fuse in (False, True):
        proc = UlyssesWanAttnProcessor(ug, fuse=fuse)
        proc.full_rotary_emb = rope
        attn.set_processor(proc)
        got = attn(x[:, sl], rotary_emb=rope[:, :, sl]).float()
        d = got - ref[:, sl]
        rel = float(d.norm() / ref[:, sl].norm())
        print(f"rank {rank} fuse={fuse}: ulysses processor vs single-GPU torch norm+rope: rel-L2 {rel:.3e}", flush=True)
        assert rel < 1e-2
    # the peer-memory data plane: projections written into the symmetric buffer, statistic pushed to the peers, rows
    # pulled by the gather kernel, output rows pushed by the attention epilogue (B = 2: one sequence after the other)
    from video_blade_b200.ulysses import UlyssesPeerPlane
    plane = UlyssesPeerPlane(ug, S // world, heads, D, dtype=torch.bfloat16, device=dev)
    proc = UlyssesWanAttnProcessor(ug, fuse=True, plane=plane)
    proc.full_rotary_emb = rope
    attn.set_processor(proc)
    for rep in range(3):                                  # repeated: buffer reuse across layers is barrier-ordered
        got = attn(x[:, sl], rotary_emb=rope[:, :, sl]).float()
        d = got - ref[:, sl]
        rel = float(d.norm() / ref[:, sl].norm())
        print(f"rank {rank} peer plane (rep {rep}): vs single-GPU torch norm+rope: rel-L2 {rel:.3e}", flush=True)
        assert rel < 1e-2
dist.destroy_process_group()
