"""Stage-by-stage CUDA-event timing of one ASA layer at a BASELINE config + full-size parity of a few heads
against a plain fp32 torch-on-GPU dense-masked evaluation.  Diagnostics; bench.py is the contract."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200.asa import AsaEngine, AsaKnobs   # noqa: E402
from oracle import asa_oracle as O                     # noqa: E402


def timeit(fn, warm=3, it=10, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def torch_ref_head(q, k, v, mask_h, kp, vp, gap, block=128):
    """fp32 reference for ONE head on the GPU: returns merged out following W:351-370 in q.dtype."""
    S, D = q.shape
    scale = 1.0 / D ** 0.5
    out = torch.empty(S, D, dtype=q.dtype, device=q.device)
    kf, vf, kpf, vpf = k.float(), v.float(), kp.float(), vp.float()
    kblk = torch.arange(S, device=q.device) // block
    for r0 in range(0, S, 2048):
        r1 = min(S, r0 + 2048)
        qf = q[r0:r1].float()
        s = (qf @ kf.T) * scale
        rb = torch.arange(r0, r1, device=q.device) // block
        s = s.masked_fill(~mask_h[rb][:, kblk], float("-inf"))
        l1 = torch.logsumexp(s, -1)
        o1 = (torch.exp(s - l1[:, None]) @ vf).to(q.dtype)
        s2 = (qf @ kpf.T) * scale
        l2 = torch.logsumexp(s2, -1)
        o2 = (torch.exp(s2 - l2[:, None]) @ vpf).to(q.dtype)
        out[r0:r1] = O.merge_lse(o1, l1[:, None].to(q.dtype), o2, l2[:, None].to(q.dtype), gap).to(q.dtype) \
            if False else _merge(o1, l1[:, None].to(q.dtype), o2, l2[:, None].to(q.dtype), gap)
    return out


def _merge(out1, lse1, out2, lse2, gap):
    lg = torch.log(torch.tensor(gap, dtype=lse1.dtype, device=lse1.device))
    lw2 = lse2 + lg
    mx = torch.maximum(lse1, lw2)
    e1 = torch.exp(lse1 - mx)
    e2 = torch.exp(lw2 - mx)
    a = e1 / (e1 + e2)
    return out1 * a + out2 * (1 - a)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="wan")
    ap.add_argument("--structured", type=float, default=0.0)
    ap.add_argument("--heads", type=int, default=0)
    ap.add_argument("--check-heads", type=int, default=1)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    if a.model == "wan":
        kn = AsaKnobs.wan()
        H, D = a.heads or 12, 128
    else:
        kn = AsaKnobs.cog()
        H, D = a.heads or 48, 64
    S = kn.width * kn.height * kn.depth + kn.text_length
    eng = AsaEngine(kn)
    t0 = time.time()
    q, k, v = O.synth_qkv(1, H, S, D, seed=0, structured=a.structured, grid=(kn.width, kn.height, kn.depth),
                          text_length=kn.text_length)
    # reference layout: [B,S,H,D] memory viewed as [B,H,S,D]
    qc, kc, vc = (x.transpose(1, 2).contiguous().cuda().transpose(1, 2) for x in (q, k, v))
    print(f"inputs {tuple(qc.shape)} in {time.time()-t0:.1f}s", flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    out, dbg = eng.forward(qc, kc, vc, return_debug=True)
    torch.cuda.synchronize()
    cnt = dbg["cnt"]
    nb = cnt.shape[-1]
    print(f"retained blocks/row: min {int(cnt.min())} mean {float(cnt.float().mean()):.2f} max {int(cnt.max())} of {nb}")
    npool = -(-S // kn.sample_gap)
    cols = cnt.clone().float() * 128
    tail = S - (nb - 1) * 128
    has_tail = dbg["mask"][..., -1]
    cols = cols - has_tail.float() * (128 - tail)
    flops = O.attention_flops(cols.cpu(), S, D, 128, npool)
    print(f"algorithmic attention FLOPs {flops/1e12:.4f} TFLOP")

    med, best = timeit(lambda: eng.forward(qc, kc, vc), it=a.iters, flush=flush)
    print(f"whole layer (blade_asa_forward): median {med:.3f} ms best {best:.3f} ms -> {flops/med/1e9:.1f} TFLOP/s sparse-eff")
    # stages
    src = eng.src_row(qc.device, S)
    med_p, _ = timeit(lambda: eng.prep(qc, kc, vc, rearrange=True), it=a.iters, flush=flush)
    (qr, kr, vr), (qm, km), (kp, vp) = eng.prep(qc, kc, vc, rearrange=True)
    med_s, _ = timeit(lambda: eng.scores_meanpool(qm, km), it=a.iters)
    sc = eng.scores_meanpool(qm, km)
    med_sel, _ = timeit(lambda: eng.select(sc, want_mask=False), it=a.iters)
    idx, cnt2, _ = eng.select(sc, want_mask=False)
    outbuf = torch.empty(1, S, H, D, dtype=qc.dtype, device="cuda").transpose(1, 2)
    med_a, best_a = timeit(lambda: eng.asa_attn(qr, kr, vr, idx, cnt2, kp, vp, out=outbuf, dst_row=src), it=a.iters,
                           flush=flush)
    print(f"prep {med_p*1e3:.1f} us | scores {med_s*1e3:.1f} us | select {med_sel*1e3:.1f} us | "
          f"attn median {med_a:.3f} ms best {best_a:.3f} ms -> {flops/med_a/1e9:.1f} TFLOP/s")
    byt = 3 * 2 * H * S * D * 2
    print(f"prep bytes (read+write q,k,v) {byt/1e6:.1f} MB -> {byt/med_p/1e6:.1f} GB/s")

    # full-size parity for a few heads
    mask = dbg["mask"]
    rr_src = src.long()
    for h in range(min(a.check_heads, H)):
        qh, kh, vh = qc[0, h][rr_src], kc[0, h][rr_src], vc[0, h][rr_src]
        ref_r = torch_ref_head(qh, kh, vh, mask[0, h], kp[0, h], vp[0, h], kn.sample_gap)
        ref = torch.empty_like(ref_r)
        ref[rr_src] = ref_r
        got = out[0, h].float()
        d = got - ref.float()
        print(f"head {h}: rel_l2 {float(d.norm()/ref.float().norm()):.3e} max_abs {float(d.abs().max()):.3e} "
              f"nan {int(torch.isnan(got).sum())}")


if __name__ == "__main__":
    main()
