"""bench.py contract (GPU) and the 8-step sampler / DiT scaffolding (CPU)."""
import json
import os
import subprocess
import sys

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
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
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
