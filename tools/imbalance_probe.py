"""How much does the static round-robin item assignment lose when rows keep different numbers of blocks?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.asa import AsaEngine, AsaKnobs
from video_blade_b200.synth import synth_qkv
kn = AsaKnobs.wan(); eng = AsaEngine(kn)
S, H, D = 32760, 12, 128
for amp in (0.0, 0.5, 0.8, 1.0, 1.3, 2.0):
    q, k, v = synth_qkv(1, H, S, D, seed=0, structured=amp, grid=(52, 30, 21))
    qc, kc, vc = (x.transpose(1, 2).contiguous().cuda().transpose(1, 2) for x in (q, k, v))
    out, cnt = eng.forward(qc, kc, vc)
    torch.cuda.synchronize()
    c = cnt.float().flatten()
    tiles = (c + 9).view(H, -1)                      # sparse + 9 pooled tiles per q-block
    pairs = tiles.view(H, -1, 2).sum(-1).flatten()   # item cost in tile-iterations (both streams)
    # static assignment: item i -> CTA i % 148 ; cost model: an item takes max(stream0, stream1) tile-steps
    t2 = tiles.view(H, -1, 2)
    item_cost = t2.max(-1).values.flatten()
    cta = torch.zeros(148)
    for i, w in enumerate(item_cost.tolist()):
        cta[i % 148] += w
    ideal = item_cost.sum() / 148
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); eng.forward(qc, kc, vc); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"amp {amp}: blocks/row min {int(c.min())} mean {c.mean():.1f} max {int(c.max())} | static makespan/ideal = "
          f"{float(cta.max()/ideal):.3f} | pair mismatch loss = {float(item_cost.sum()*2/tiles.sum()):.3f} | layer {sorted(ts)[2]:.3f} ms")
