"""A/B of the half-tile items (KV list of the last query tiles split across two CTAs): attention-kernel time at the
Wan size for 12 heads (one GPU) and 3 heads (one rank of ulysses 4), uniform and non-uniform rows.
    python tools/ab_xsplit.py                      # default (half tiles on)
    BLADE_NO_XSPLIT=1 python tools/ab_xsplit.py    # off; the outputs of the two runs are compared through /tmp"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.asa import AsaEngine, AsaKnobs
tag = "off" if os.environ.get("BLADE_NO_XSPLIT") == "1" else "on"
knobs = " ".join(f"{k[6:]}={v}" for k, v in os.environ.items() if k in ("BLADE_XSPLIT_PAIRS", "BLADE_SOLO_PAIRS"))
S, D, nb = 32760, 128, 256
for H in [int(x) for x in os.environ.get('AB_HEADS', '12,3').split(',')]:
    eng = AsaEngine(AsaKnobs.wan(use_rearrange=False))
    torch.manual_seed(0)
    q, k, v = (torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16).transpose(1, 2) for _ in range(3))
    kp = torch.randn(1, H, 1092, D, device="cuda", dtype=torch.bfloat16)
    vp = torch.randn(1, H, 1092, D, device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for label, counts in (("uniform 43", torch.full((1, H, nb), 43, device="cuda")),
                          ("random 12..43", torch.randint(12, 44, (1, H, nb), device="cuda"))):
        score = torch.rand(1, H, nb, nb, device="cuda")
        kth = torch.sort(score, dim=-1, descending=True).values.gather(-1, (counts - 1).clamp(min=0)[..., None])
        idx, cnt = eng.mask_to_index(score >= kth)
        for _ in range(3):
            out = eng.asa_attn(q, k, v, idx, cnt, kp, vp)
        ts = []
        for _ in range(15):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); out = eng.asa_attn(q, k, v, idx, cnt, kp, vp); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        o = out[0] if isinstance(out, tuple) else out
        f = f"/tmp/xsplit_{H}_{label.split()[0]}"
        msg = ""
        other = f + ("_on" if tag == "off" else "_off") + ".pt"
        torch.save(o.float().cpu(), f + f"_{tag}.pt")
        if os.path.exists(other):
            ref = torch.load(other)
            d = o.float().cpu() - ref
            msg = f"  vs other setting: rel_l2 {float(d.norm() / ref.norm()):.2e} max_abs {float(d.abs().max()):.2e}"
        print(f"xsplit {tag:3s} {knobs:36s} H={H:2d} {label:14s}: median {ts[7]:.4f} ms  min {ts[0]:.4f}{msg}", flush=True)
