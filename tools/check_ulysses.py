"""torchrun --nproc-per-node P tools/check_ulysses.py : the Ulysses path (fused scatter + ASA on H/P heads + gather)
must equal the single-GPU layer on the full tensor, bit for bit (ASA is independent per head)."""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.asa import AsaEngine, AsaKnobs
from video_blade_b200.ulysses import UlyssesGroup
from oracle.asa_oracle import synth_qkv

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
grid = (26, 16, 4)
S, H, D = grid[0] * grid[1] * grid[2], 4 * world, 128
assert S % world == 0
kn = AsaKnobs.wan(width=grid[0], height=grid[1], depth=grid[2], max_retain_ratio=0.4)
eng = AsaEngine(kn)
q, k, v = synth_qkv(1, H, S, D, seed=5, structured=2.0, grid=grid)
qf, kf, vf = (x.transpose(1, 2).contiguous().cuda() for x in (q, k, v))          # [1,S,H,D]
ref, _ = eng.forward(qf.transpose(1, 2), kf.transpose(1, 2), vf.transpose(1, 2))    # [1,H,S,D]
ug = UlyssesGroup(world, rank, world)
sl = slice(rank * (S // world), (rank + 1) * (S // world))
for fused in (False, True):
    if fused:
        gq, gk, gv, vrow, keep = ug.scatter_heads_fused(qf[:, sl].contiguous(), kf[:, sl].contiguous(), vf[:, sl].contiguous())
        o, _ = eng.forward(gq, gk, gv, virtual_rows=vrow)
    else:
        gq, gk, gv = ug.scatter_heads(qf[:, sl].contiguous(), kf[:, sl].contiguous(), vf[:, sl].contiguous())
        o, _ = eng.forward(gq.transpose(1, 2), gk.transpose(1, 2), gv.transpose(1, 2))
    mine = ug.gather_heads(o.transpose(1, 2))                                      # [1,S/P,H,D]
    want = ref.transpose(1, 2)[:, sl]
    ok = torch.equal(mine, want)
    print(f"rank {rank} fused={fused}: ulysses == single-GPU: {ok}  max|d| {float((mine.float()-want.float()).abs().max()):.3e}", flush=True)
    assert ok
dist.destroy_process_group()
