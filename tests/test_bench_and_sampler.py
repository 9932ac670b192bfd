"""bench.py contract (GPU) and the 8-step sampler / DiT scaffolding (CPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_generate_new_matches_reference_update_rule():
    """train_wanx_tdm.py:1402-1443 restated: with v == const the K-step recursion has a closed form per step."""
    from video_blade_b200.dit import flow_sigma, generate_new
    torch.manual_seed(0)
    x = torch.randn(2, 4, 2, 4, 4)
    calls = []

    def vel(x_t, T):
        calls.append(T.clone())
        return 0.3 * x_t

    out = generate_new(vel, x, steps=4, eta=1.0, flow_shift=3.0)
    assert [int(t[0]) for t in calls] == [999, 749, 499, 249]                  # T <- T - 1000/K  (TW:1435)
    # manual recursion
    x_t, T = x, torch.full((2,), 999)
    for _ in range(4):
        s = flow_sigma(T).view(2, 1, 1, 1, 1)
        v = 0.3 * x_t
        x0 = x_t - s * v                                                        # TW:1426
        eps = x_t + (1 - s) * v                                                 # TW:1431
        T = T - 250
        s2 = flow_sigma(T.clamp(min=0)).view(2, 1, 1, 1, 1)
        x_t = (1 - s2) * x0 + s2 * eps                                          # add_noise, TW:1437
    assert torch.allclose(out, x0)
    assert abs(float(flow_sigma(torch.tensor([500]))[0]) - 3 * 0.5 / (1 + 2 * 0.5)) < 1e-6


def test_dit_scaffolds_shapes_and_cfg():
    from video_blade_b200.dit import CogLikeDiT, WanLikeDiT, make_velocity_fn
    from video_blade_b200.modify_cogvideo import SageAttnCogVideoXAttnProcessor
    from video_blade_b200.modify_wan import WanAttnProcessor2_0

    class Dense(torch.nn.Module):
        def forward(self, q, k, v, **kw):
            return torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.manual_seed(0)
    wan = WanLikeDiT(dim=64, heads=2, ffn=128, layers=2, text_dim=32).eval()
    for b in wan.blocks:
        b.attn1.inner_attention = Dense()
        b.attn1.set_processor(WanAttnProcessor2_0())
    lat = torch.randn(1, 16, 3, 8, 12)
    p, n = torch.randn(1, 7, 32), torch.randn(1, 7, 32)
    with torch.no_grad():
        v1 = make_velocity_fn(wan, p, n, 1.0)(lat, torch.tensor([999]))
        v5 = make_velocity_fn(wan, p, n, 5.0)(lat, torch.tensor([999]))
        vc, vu = wan(lat, torch.tensor([999]), p), wan(lat, torch.tensor([999]), n)
    assert v1.shape == lat.shape and torch.allclose(v1, vc, atol=1e-5)
    assert torch.allclose(v5, vu + 5.0 * (vc - vu), atol=1e-4)                  # classifier-free guidance
    cog = CogLikeDiT(dim=64, heads=2, layers=2, text_dim=32, temb_dim=16).eval()
    for i, b in enumerate(cog.transformer_blocks):
        b.attn1.inner_attention = Dense()
        b.attn1.set_processor(SageAttnCogVideoXAttnProcessor(i))
    with torch.no_grad():
        out = cog(torch.randn(1, 3, 16, 8, 12), torch.tensor([500]), torch.randn(1, 5, 32))
    assert out.shape == (1, 3, 16, 8, 12) and torch.isfinite(out).all()


@pytest.mark.gpu
def test_bench_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3",
                        "--no-cpu-baseline", "--no-clip"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["higher_is_better"] is True
    assert "workload" in line["config"] and "model" not in line["config"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"]
    assert line["e2e"]["h2d_bytes_per_step"] == 3 * 32760 * 12 * 128 * 2
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"]
    assert line["roofline"]["bound"] == "tensor" and 0.2 < line["roofline"]["frac"] < 1.2
    assert line["gpu_launches"] == 3 * 5 and line["value"] > 100
    mg = line["roofline_maskgen"]                                 # 207.6 MB over the WHOLE mask-gen chain
    assert mg["bound"] == "hbm" and mg["bytes"] == 207581184 and 0.02 < mg["frac"] < 1.0
    assert abs(mg["chain_ms"] - sum(line["config"]["stage_ms"][k] for k in ("prep", "scores", "select"))) < 1e-6
    assert line["hoisted"]["bit_equal_to_per_layer_gather"] is True and "pool" in line["config"]["stage_ms"]
    assert abs(line["config"]["algorithmic_tflop_per_step"] - 1.327) < 0.01


def test_reference_arm_json_contract_cpu():
    """`bench.py --impl reference` needs no GPU: the oracle port on the host cores, one bounded sample per step,
    same metric / unit / config as the GPU arm, `e2e` repeating the line's own value with zero transfer bytes."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "TFLOP/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("ASA sparse-effective attention throughput")
    assert "workload" in line["config"] and "32760 tok" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "heads" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and 0 < line["value"] < 5


def test_algorithmic_flops_match_baseline_formula():
    """BASELINE.md section 3 at the headline shape: Wan 32 760 tokens, 12 heads, d = 128, 43 of 256 blocks per row
    (the ragged last block selected on a third of the rows), 1 092 pooled keys."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    S, D, H, nb, n_pool = 32760, 128, 12, 256, 1092
    cnt = torch.full((1, H, nb), 43)
    last = torch.zeros(1, H, nb, dtype=torch.bool)
    last[..., ::3] = True
    got = bench.algorithmic_flops(cnt, last, S, D, n_pool)
    rows = torch.full((nb,), 128.0, dtype=torch.float64)
    rows[-1] = S - 255 * 128                                     # 120
    cols = torch.full((nb,), 43 * 128.0, dtype=torch.float64)
    cols[::3] -= 8                                               # the last key block holds 120 real keys
    want = 4 * D * H * float((rows * (cols + n_pool)).sum())
    assert abs(got - want) / want < 1e-12
    assert abs(got / 1e12 - 1.327) < 0.005


# ------------------------------------------------------------------ samplers of record
def test_unipc_flow_schedule_and_first_order_equals_generate_new():
    """(a) the sigma / timestep schedule of UniPCMultistepScheduler(use_flow_sigmas, flow_shift=3) in closed form;
    (b) its first-order predictor step is exactly generate_new's update with eta = 1 (TW:1426-1437) for the same pair
    of sigmas -- both are the exact solution of the flow ODE for a constant x0 prediction."""
    from video_blade_b200.dit import flow_sigma
    from video_blade_b200.samplers import UniPCFlowScheduler, flow_sigmas
    sig = flow_sigmas(8, 3.0)
    assert sig.shape == (9,) and sig[-1] == 0.0 and all(sig[i] > sig[i + 1] for i in range(8))
    raw = 1.0 - np.linspace(1.0, 1e-3, 9)[::-1][:-1]
    assert np.allclose(sig[:-1], 3.0 * raw / (1.0 + 2.0 * raw))
    # generate_new walks the integer timesteps 999, 874, ...: same schedule up to the rounding of the timestep
    gn = flow_sigma(torch.tensor([999 - 125 * i for i in range(8)]), 3.0).numpy()
    assert np.abs(gn - sig[:-1]).max() < 2e-3
    sch = UniPCFlowScheduler(solver_order=1, use_corrector=False)
    sch.set_timesteps(8)
    assert [int(t) for t in sch.timesteps][:3] == [999, 954, 899]          # int(sigma * 1000), like diffusers
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 4, generator=g, dtype=torch.float64)
    for i in range(8):
        v = torch.randn(2, 3, 4, generator=g, dtype=torch.float64)
        s, s2 = sch.sigmas[i], sch.sigmas[i + 1]
        x0, eps = x - s * v, x + (1 - s) * v                                  # TW:1426, 1431
        want = (1 - s2) * x0 + s2 * eps                                       # TW:1437 with eta = 1
        got = sch.step(v, x)
        assert torch.allclose(got, want, rtol=1e-10, atol=1e-12), i
        x = got
    assert torch.allclose(x, x0)                                              # final sigma 0: the sample is the x0 prediction


def test_unipc_flow_second_order_beats_first_order_and_is_exact_for_constant_x0():
    from video_blade_b200.samplers import sample_unipc_flow
    g = torch.Generator().manual_seed(1)
    x1 = torch.randn(4, 5, generator=g, dtype=torch.float64)                 # "data"
    noise = torch.randn(4, 5, generator=g, dtype=torch.float64)
    # (1) a model whose x0 prediction is constant: every order reproduces x1 exactly

    class ConstX0:
        def __init__(self, sch_sigmas):
            self.s, self.i = sch_sigmas, 0

        def __call__(self, x, T):
            sig = self.s[self.i]
            self.i += 1
            return (x - x1) / sig                                             # v such that x - sigma v == x1
    from video_blade_b200.samplers import UniPCFlowScheduler
    sch = UniPCFlowScheduler()
    sch.set_timesteps(8)
    out = sample_unipc_flow(ConstX0(sch.sigmas), noise, steps=8)
    assert torch.allclose(out, x1, atol=1e-9)

    # (2) a smooth non-constant x0 prediction x0(x, sigma) = x1 + 0.3 sigma sin(x), integrated from sigma_0 down to the
    # last non-zero sigma of the 8-step schedule (the final jump to sigma = 0 returns the x0 prediction for every
    # order).  Reference: the same sigma range in 250 first-order sub-steps per interval.
    coarse = UniPCFlowScheduler().sigmas.clone()

    def run(sigmas, n_steps, order, corr):
        sch = UniPCFlowScheduler(solver_order=order, use_corrector=corr)
        sch.set_timesteps(len(sigmas) - 1)
        sch.sigmas = sigmas
        x = noise.clone()
        for i in range(n_steps):
            sig = sch.sigmas[i]
            x0 = x1 + 0.3 * sig * torch.sin(x)
            x = sch.step((x - x0) / sig, x)
        return x
    def refine(kk):
        return torch.cat([torch.linspace(float(coarse[i]), float(coarse[i + 1]), kk + 1, dtype=torch.float64)[:-1]
                          for i in range(7)] + [coarse[7:]])
    ref = run(refine(2000), 7 * 2000, 1, False)
    err = {(o, kk): float((run(refine(kk), 7 * kk, o, o == 2) - ref).abs().max()) for o in (1, 2) for kk in (1, 2, 4)}
    assert err[(2, 1)] < 0.5 * err[(1, 1)], err                   # at the 8-step schedule itself
    assert 1.6 < err[(1, 2)] / err[(1, 4)] < 2.6, err             # first order: halving the step halves the error
    assert err[(2, 2)] / err[(2, 4)] > 3.0, err                   # predictor-corrector order 2: at least quarters it


def test_cogvideox_dpm_trailing_schedule_and_marginals():
    """(a) "trailing" timesteps 999, 874, ... and the zero-terminal-SNR alphas; (b) every step of the SDE DPM-Solver++
    maps sqrt(a) x0 + sqrt(1-a) eps to a sample with mean sqrt(a') x0 and variance 1-a' when the x0 prediction is exact
    -- the property that fixes mult1 / mult2 / mult_noise; (c) the last step returns the x0 prediction."""
    from video_blade_b200.samplers import CogVideoXDPMScheduler, trailing_timesteps
    assert trailing_timesteps(8).tolist() == [999, 874, 749, 624, 499, 374, 249, 124]
    sch = CogVideoXDPMScheduler()
    ac = sch.alphas_cumprod
    assert float(ac[-1]) == 0.0 and abs(float(ac[0]) - (1 - 0.00085)) < 1e-12 and bool((ac[:-1] > ac[1:]).all())
    g = torch.Generator().manual_seed(2)
    x0 = torch.randn(3, 7, generator=g, dtype=torch.float64)
    ts = [int(t) for t in sch.timesteps]
    for i, t in enumerate(ts):
        a_t = ac[t]
        a_p = ac[t - 125] if t - 125 >= 0 else torch.tensor(1.0, dtype=torch.float64)
        eps = torch.randn(3, 7, generator=g, dtype=torch.float64)
        x = a_t.sqrt() * x0 + (1 - a_t).sqrt() * eps
        v = a_t.sqrt() * eps - (1 - a_t).sqrt() * x0                          # the exact v-prediction
        # deterministic skeleton (noise_fn None): mean part
        prev, pred = sch.step(v, x0 if i else None, t, ts[i - 1] if i else None, x)
        assert torch.allclose(pred, x0, atol=1e-9)
        h = sch._lam(a_p) - sch._lam(a_t)
        m1 = ((1 - a_p) / (1 - a_t)).sqrt() * torch.exp(-h)
        m_noise = (1 - a_p).sqrt() * (1 - torch.exp(-2 * h)).sqrt()
        want_mean = a_p.sqrt() * x0 + m1 * (1 - a_t).sqrt() * eps
        assert torch.allclose(prev, want_mean, atol=1e-9), t
        assert abs(float(m1 ** 2 * (1 - a_t) + m_noise ** 2 - (1 - a_p))) < 1e-12   # total noise variance = 1 - a'
    assert torch.allclose(prev, x0, atol=1e-9)


def test_clip_samplers_run_on_the_scaffold():
    """The three samplers drive the same guided-velocity callable (ASA untouched): shapes and finiteness on a tiny DiT."""
    from video_blade_b200.dit import WanLikeDiT, generate_new, make_velocity_fn
    from video_blade_b200.modify_wan import WanAttnProcessor2_0
    from video_blade_b200.samplers import sample_cogvideox_dpm, sample_unipc_flow

    class Dense(torch.nn.Module):
        def forward(self, q, k, v, **kw):
            return torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.manual_seed(0)
    wan = WanLikeDiT(dim=64, heads=2, ffn=128, layers=1, text_dim=32).eval()
    for b in wan.blocks:
        b.attn1.inner_attention = Dense()
        b.attn1.set_processor(WanAttnProcessor2_0())
    noise = torch.randn(1, 16, 3, 8, 12)
    vel = make_velocity_fn(wan, torch.randn(1, 7, 32), torch.randn(1, 7, 32), 5.0)
    with torch.no_grad():
        a = generate_new(vel, noise, steps=4)
        b = sample_unipc_flow(vel, noise, steps=4)
        c = sample_cogvideox_dpm(vel, noise, steps=4, generator=torch.Generator().manual_seed(0))
    for o in (a, b, c):
        assert o.shape == noise.shape and torch.isfinite(o).all()


def test_hoisted_gilbert_permutation_is_equivalent_on_the_scaffold():
    """permuting the tokens (and their rotary table) into curve order once per forward and running the
    attention without its per-layer gather gives the same model output -- every other op is token-wise."""
    from video_blade_b200 import wanx_blocksparseattn as W
    from video_blade_b200.dit import WanLikeDiT
    from video_blade_b200.modify_wan import WanAttnProcessor2_0

    class CurveLocalInner(torch.nn.Module):
        """Stand-in for ASA on CPU with the property that matters here: it is NOT permutation equivariant -- like ASA it
        gathers into curve order (when use_rearrange), attends within blocks of 16 consecutive curve positions, and
        scatters back."""
        use_rearrange = True

        def forward(self, q, k, v, **kw):
            from video_blade_b200.asa import token_order
            S = q.shape[2]
            order = torch.from_numpy(token_order(W._knobs())).long()
            if self.use_rearrange:
                q, k, v = q[:, :, order], k[:, :, order], v[:, :, order]
            o = torch.cat([torch.nn.functional.scaled_dot_product_attention(q[:, :, i:i + 16], k[:, :, i:i + 16],
                                                                             v[:, :, i:i + 16]) for i in range(0, S, 16)], 2)
            if self.use_rearrange:
                back = torch.empty_like(o)
                back[:, :, order] = o
                o = back
            return o
    old = (W.width, W.height, W.depth)
    W.width, W.height, W.depth = 6, 4, 3                                       # 72 tokens: lat [1,16,3,8,12], patch (1,2,2)
    try:
        torch.manual_seed(0)
        net = WanLikeDiT(dim=64, heads=2, ffn=128, layers=2, text_dim=32).eval()
        inner = CurveLocalInner()
        for b in net.blocks:
            b.attn1.inner_attention = inner
            b.attn1.set_processor(WanAttnProcessor2_0())
        lat, ctx, t = torch.randn(1, 16, 3, 8, 12), torch.randn(1, 7, 32), torch.tensor([500])
        with torch.no_grad():
            ref = net(lat, t, ctx)
            net.set_hoisted_permutation(True)
            assert inner.use_rearrange is False
            got = net(lat, t, ctx)
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), float((got - ref).abs().max())
    finally:
        W.width, W.height, W.depth = old
