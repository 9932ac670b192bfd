"""GPU bring-up diagnostics (run on the B200 box): probes the tcgen05 descriptors / TMEM layouts through the
C ABI, then the small kernels, then one small attention call, printing enough detail to localise a failure.
Writes gpurun_out/bringup.log.  Test infrastructure only."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200 import _lib                      # noqa: E402
from video_blade_b200.asa import AsaEngine, AsaKnobs   # noqa: E402
from oracle import asa_oracle as O                     # noqa: E402  (checker only)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "bringup.log"), "a")


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def describe(name, got, want):
    got = got.float().cpu()
    want = want.float().cpu()
    err = (got - want).abs()
    rel = (got - want).norm() / want.norm().clamp_min(1e-20)
    say(f"  {name}: max_abs={err.max():.4e} rel_l2={rel:.4e} nan={int(torch.isnan(got).sum())} "
        f"got_absmean={got.abs().mean():.4e} want_absmean={want.abs().mean():.4e}")
    return float(rel)


def probe(D):
    lib = _lib.load()
    g = torch.Generator().manual_seed(D)
    q = torch.randn(128, D, generator=g).bfloat16().cuda()
    k = torch.randn(128, D, generator=g).bfloat16().cuda()
    s = torch.zeros(128, 128, device="cuda")
    _lib.check(lib.blade_probe_qk(q.data_ptr(), k.data_ptr(), s.data_ptr(), D, _lib.current_stream()))
    torch.cuda.synchronize()
    want = q.float() @ k.float().T
    say(f"probe QK^T D={D}")
    r = describe("S", s, want)
    if r > 1e-3:
        describe("S vs want^T", s, want.T)
        bad = ((s - want).abs() > 1e-2 * want.abs().max())
        say("   bad rows:", bad.any(1).sum().item(), "bad cols:", bad.any(0).sum().item(),
            "first bad cols:", bad.any(0).nonzero().flatten()[:16].tolist(),
            "first bad rows:", bad.any(1).nonzero().flatten()[:16].tolist())
    p = torch.rand(128, 128, generator=g).cuda()
    v = torch.randn(128, D, generator=g).bfloat16().cuda()
    o = torch.zeros(128, D, device="cuda")
    _lib.check(lib.blade_probe_pv(p.data_ptr(), v.data_ptr(), o.data_ptr(), D, _lib.current_stream()))
    torch.cuda.synchronize()
    want = p.bfloat16().float() @ v.float()
    say(f"probe PV D={D}")
    r2 = describe("O", o, want)
    if r2 > 1e-3:
        bad = ((o - want).abs() > 1e-2 * want.abs().max())
        say("   bad rows:", bad.any(1).sum().item(), "bad cols:", bad.any(0).sum().item(),
            "first bad cols:", bad.any(0).nonzero().flatten()[:16].tolist())
        # diagnose K-order problems: does O match P[:, perm] @ V for simple permutations?
        pe = p.bfloat16().float()
        sw = pe.view(128, 64, 2).flip(-1).reshape(128, 128)
        describe("O vs pairswapped-P @ V", o, sw @ v.float())
    return r, r2


def small_kernels():
    eng = AsaEngine(AsaKnobs.wan(width=26, height=15, depth=4))
    for nb in (61, 139, 256):
        g = torch.Generator().manual_seed(nb)
        sc = torch.softmax(torch.randn(1, 2, nb, nb, generator=g) * 2.5, -1)
        lo, hi = O.retain_bounds(nb, 0.05, 0.17, "wan")
        want, _ = O.select_blocks_energy(sc, lo, hi, 0.95)
        idx, cnt, mask = eng.select(sc.cuda(), lo=lo, hi=hi, force_last=0)
        torch.cuda.synchronize()
        widx, wcnt = O.mask_to_index_list(want)
        say(f"select nb={nb}: mask_equal={torch.equal(mask.cpu(), want)} idx_equal={torch.equal(idx.cpu(), widx)} "
            f"cnt_equal={torch.equal(cnt.cpu(), wcnt)}")
    S = 26 * 15 * 4
    q, k, v = O.synth_qkv(1, 2, S, 128, seed=1)
    (qr, kr, vr), (qm, km), (kp, vp) = eng.prep(q.cuda(), k.cuda(), v.cuda(), rearrange=True)
    torch.cuda.synchronize()
    rr = O.GilbertRearranger(26, 15, 4)
    say("prep:")
    say("  q_r equal:", torch.equal(qr.cpu(), rr.rearrange(q)), " v_r equal:", torch.equal(vr.cpu(), rr.rearrange(v)))
    qp = O.pad_to_multiple(rr.rearrange(q), 128).float()
    describe("q_mean", qm, qp.reshape(1, 2, -1, 128, 128).mean(3))
    describe("k_pool", kp, O.simple_pooling(rr.rearrange(k), 30))
    sc = eng.scores_meanpool(qm, km)
    torch.cuda.synchronize()
    describe("scores", sc, O.estimator_meanpool(rr.rearrange(q), rr.rearrange(k), 128))


def small_attn(D, H, S, with_pool, seed=0):
    eng = AsaKnobs.wan(use_rearrange=False, sample_gap=30)
    eng = AsaEngine(eng)
    q, k, v = O.synth_qkv(1, H, S, D, seed=seed)
    nb = -(-S // 128)
    g = torch.Generator().manual_seed(seed + 5)
    mask = torch.rand(1, H, nb, nb, generator=g) < 0.35
    mask[..., 0] = True
    mask[:, :, -1, -1] = True
    idx, cnt = O.mask_to_index_list(mask)
    qc, kc, vc = q.cuda(), k.cuda(), v.cuda()
    if not with_pool:
        out, lse = eng.block_sparse_attn(qc, kc, vc, idx.cuda(), cnt.cuda())
        torch.cuda.synchronize()
        wout, wlse = O.dense_masked_attention(q, k, v, mask)
        say(f"block_sparse_attn D={D} H={H} S={S}")
        r = describe("out", out, wout)
        describe("lse", lse, wlse)
        return r
    kp = O.simple_pooling(k, 30)
    vp = O.simple_pooling(v, 30)
    out = eng.asa_attn(qc, kc, vc, idx.cuda(), cnt.cuda(), kp.cuda(), vp.cuda())
    torch.cuda.synchronize()
    o1, l1 = O.dense_masked_attention(q, k, v, mask)
    o2, l2 = O.standard_attn(q, kp, vp)
    want = O.merge_lse(o1, l1.unsqueeze(-1).to(q.dtype), o2, l2.unsqueeze(-1).to(q.dtype), 30)
    say(f"asa_attn (pooled+merge) D={D} H={H} S={S}")
    return describe("out", out, want)


def main():
    say("=== bring-up on", torch.cuda.get_device_name(0))
    _lib.check(_lib.load().blade_device_check())
    steps = [("probe128", lambda: probe(128)), ("probe64", lambda: probe(64)), ("small", small_kernels),
             ("attn128", lambda: small_attn(128, 2, 1560, False)), ("attn64", lambda: small_attn(64, 3, 940, False)),
             ("asa128", lambda: small_attn(128, 2, 1560, True)), ("asa64", lambda: small_attn(64, 3, 940, True))]
    only = sys.argv[1:]
    for name, fn in steps:
        if only and name not in only:
            continue
        try:
            fn()
        except Exception:
            say(f"!! step {name} raised:\n{traceback.format_exc()}")
            break


if __name__ == "__main__":
    main()
