"""GPU diagnostic: sampled-max estimator kernel vs the oracle restatement (same offsets)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.asa import AsaEngine, AsaKnobs
from oracle import asa_oracle as O
for (grid, H, D, T, flavor) in [((26, 15, 4), 2, 128, 0, "wan"), ((15, 10, 6), 3, 64, 40, "cog"), ((52, 30, 5), 2, 128, 0, "wan")]:
    S = grid[0] * grid[1] * grid[2] + T
    q, k, v = O.synth_qkv(1, H, S, D, seed=3, structured=2.0, grid=grid, text_length=T)
    kn = (AsaKnobs.cog if flavor == "cog" else AsaKnobs.wan)(width=grid[0], height=grid[1], depth=grid[2], text_length=T, estimator="sampled_max", max_retain_ratio=0.3)
    eng = AsaEngine(kn)
    g = torch.Generator().manual_seed(1)
    qo = O.draw_sample_offsets(1, H, 128, 32, g); ko = O.draw_sample_offsets(1, H, 128, 32, g)
    want = O.estimator_sampled_max(q, k, 128, qo, ko).float()
    got = eng.scores_sampled(q.cuda(), k.cuda(), qo.cuda(), ko.cuda()).cpu()
    torch.cuda.synchronize()
    eq = (got == want).float().mean().item()
    rel = ((got - want).abs() / want.clamp_min(1e-9)).max().item()
    print(f"{flavor} S={S} D={D}: exact-equal fraction {eq:.4f}  max rel diff {rel:.3e}  nan {int(torch.isnan(got).sum())}  rowsum {got.sum(-1).mean():.4f}")
    cfg = O.ASAConfig(flavor=flavor, width=grid[0], height=grid[1], depth=grid[2], text_length=T, max_retain_ratio=0.3,
                      sample_gap=kn.sample_gap, estimator="sampled_max")
    mg, _ = O.select_mask(got, cfg); mw, _ = O.select_mask(want, cfg)
    print("   mask IoU", ((mg & mw).sum() / (mg | mw).sum()).item())
    out, dbg = eng.forward(q.cuda(), k.cuda(), v.cuda(), return_debug=True, sample_offsets=(qo.cuda(), ko.cuda()))
    ref = O.asa_forward(q, k, v, cfg, qo, ko)
    d = out.float().cpu() - ref.out.float()
    print("   layer vs oracle: rel_l2", float(d.norm() / ref.out.float().norm()), "max_abs", float(d.abs().max()),
          "mask equal", torch.equal(dbg["mask"].cpu(), ref.mask))
# timing at Wan size
kn = AsaKnobs.wan(estimator="sampled_max"); eng = AsaEngine(kn)
S, H, D = 32760, 12, 128
q, k, v = (torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16).transpose(1, 2) for _ in range(3))
qo = eng.draw_offsets(1, H, q.device); ko = eng.draw_offsets(1, H, q.device)
for _ in range(3): sc = eng.scores_sampled(q, k, qo, ko)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
for _ in range(10): sc = eng.scores_sampled(q, k, qo, ko)
e1.record(); torch.cuda.synchronize()
print(f"Wan-size sampled estimator (sample + score): {e0.elapsed_time(e1)/10*1e3:.1f} us  (QK^T 0.206 TFLOP)")
for _ in range(3): out, cnt = eng.forward(q, k, v)
torch.cuda.synchronize(); e0.record()
for _ in range(10): out, cnt = eng.forward(q, k, v)
e1.record(); torch.cuda.synchronize()
print(f"Wan-size whole layer with sampled_max estimator: {e0.elapsed_time(e1)/10:.3f} ms, retained mean {float(cnt.float().mean()):.1f}")
