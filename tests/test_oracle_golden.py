"""CPU: the oracle restatement against fixtures produced by the reference's own code
(oracle/make_golden.py, run in the build container against /root/reference)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import asa_oracle as O
from oracle.gilbert import gilbert3d_order, gilbert_permutations
from conftest import GOLDEN, bf16_from_bits, load_npz


def test_gilbert_matches_reference_hashes():
    with open(os.path.join(GOLDEN, "gilbert_hashes.json")) as f:
        gold = json.load(f)
    assert len(gold) >= 10
    for key, g in gold.items():
        w, h, d = map(int, key.split("x"))
        c2r, r2c = gilbert_permutations(w, h, d)
        assert c2r[:16].tolist() == g["head"][: len(c2r[:16])], key
        assert hashlib.sha256(c2r.astype(np.int64).tobytes()).hexdigest() == g["sha256_curve2raster"], key
        assert hashlib.sha256(r2c.astype(np.int64).tobytes()).hexdigest() == g["sha256_raster2curve"], key


@pytest.mark.parametrize("grid", [(8, 6, 4), (2, 2, 2), (3, 5, 7), (45, 30, 13), (52, 30, 21)])
def test_gilbert_properties(grid):
    # the properties the reference's own script checks (test_gilbert_rearranger.py:70-262, 264-309)
    w, h, d = grid
    xyz = gilbert3d_order(w, h, d)
    n = w * h * d
    c2r, r2c = gilbert_permutations(w, h, d)
    assert sorted(c2r.tolist()) == list(range(n))                      # bijection / index range
    assert np.array_equal(c2r[r2c], np.arange(n))                      # inverse consistency
    assert np.array_equal(r2c[c2r], np.arange(n))
    step = np.abs(np.diff(xyz, axis=0)).sum(1)
    assert step.max() <= 2 and (step == 1).mean() > 0.9                # (almost) unit-step; odd sizes add diagonals


def test_gilbert_text_at_tail_roundtrip():
    # test_gilbert_rearranger.py:76-80,137-140: 8x6x4 grid, text 10, seed 42
    torch.manual_seed(42)
    rr = O.GilbertRearranger(8, 6, 4, text_length=10)
    x = torch.randn(1, 2, 10 + 8 * 6 * 4, 8)
    y = rr.rearrange(x)
    assert torch.equal(y[..., -10:, :], x[..., :10, :])                # text preserved, at the tail
    assert torch.equal(rr.reversed_rearrange(y), x)                    # round trip


def test_helpers_match_reference():
    z, _ = load_npz("helpers.npz")
    x = bf16_from_bits(z["x"])
    assert torch.equal(O.pad_to_multiple(x, 128), bf16_from_bits(z["pad128"]))
    assert torch.equal(O.simple_pooling(x, 30), bf16_from_bits(z["pool30"]))
    assert torch.equal(O.simple_pooling(x, 15), bf16_from_bits(z["pool15"]))
    g = torch.Generator().manual_seed(11)
    off = O.draw_sample_offsets(1, 2, 128, 32, g)
    got = O.sample_tokens(O.pad_to_multiple(x, 128), 128, off)
    assert torch.equal(got, bf16_from_bits(z["sampled"]))


def test_estimator_matches_triton_kernel():
    z = np.load(os.path.join(GOLDEN, "estimator_triton_fp32.npz"))
    sq, sk, po = (torch.from_numpy(z[k]) for k in ("sq", "sk", "po"))
    nk = 32
    ident = torch.arange(nk).view(1, 1, nk).expand(1, 2, nk)
    got = O.estimator_sampled_max(sq, sk, nk, ident, ident)
    assert got.shape == po.shape
    # fp32 dot-product association differs between torch.matmul and the Triton interpreter
    assert torch.allclose(got, po, rtol=2e-5, atol=1e-7)


def test_select_matches_reference_bit_exact():
    z, meta = load_npz("select_cases.npz")
    assert len(meta) == 32
    seen = set()
    for m in meta:
        c = m["case"]
        sc = torch.from_numpy(z[f"scores_{c}"])
        nb = m["nb"]
        want = torch.from_numpy(np.unpackbits(z[f"mask_{c}"], axis=-1)[..., :nb].astype(bool))
        lo, hi = O.retain_bounds(nb, m["min_ratio"], m["max_ratio"], m["flavor"])
        got, k = O.select_blocks_energy(sc, lo, hi, m["thr"], force_last=2 if m["flavor"] == "cog" else 0)
        assert torch.equal(got, want), m
        seen.add((int(k.min()) == lo, int(k.max()) == hi))
    assert (True, True) in seen or len(seen) > 1      # both clamps exercised somewhere


def test_retain_bounds_match_survey():
    assert O.retain_bounds(256, 0.05, 0.17, "wan") == (12, 43)
    assert O.retain_bounds(139, 0.05, 0.1, "cog") == (6, 13)
    assert O.retain_bounds(122, 0.05, 0.17, "wan") == (6, 20)
    assert O.retain_bounds(61, 0.05, 0.17, "wan") == (3, 10)


@pytest.mark.parametrize("name", ["layer_wan_small.npz", "layer_cog_small.npz"])
def test_full_layer_matches_reference_forward(name):
    """AdaptiveBlockSparseAttnTrain.forward (reference code, external kernel substituted) vs oracle."""
    z, m = load_npz(name)
    q, k, v = (bf16_from_bits(z[n]) for n in ("q", "k", "v"))
    want = bf16_from_bits(z["out"])
    want_nr = bf16_from_bits(z["out_norearrange"])
    flavor = "cog" if m["text_length"] else "wan"
    w, h, d = m["grid"]
    cfg = O.ASAConfig(flavor=flavor, width=w, height=h, depth=d, text_length=m["text_length"],
                      sample_gap=m["sample_gap"], max_retain_ratio=m["max_retain_ratio"],
                      min_retain_ratio=m["min_retain_ratio"], estimator="sampled_max")
    B, H = q.shape[:2]
    for use_rr, ref in ((True, want), (False, want_nr)):
        cfg.use_rearrange = use_rr
        g = torch.Generator().manual_seed(m["rng_seed"])
        qo = O.draw_sample_offsets(B, H, cfg.block_size, cfg.num_keep, g)
        ko = O.draw_sample_offsets(B, H, cfg.block_size, cfg.num_keep, g)
        res = O.asa_forward(q, k, v, cfg, qo, ko)
        assert torch.equal(res.out, ref), (name, use_rr, (res.out.float() - ref.float()).abs().max())
        # the gather evaluation is the same function within fp32 summation-order noise
        res_fast = O.asa_forward(q, k, v, cfg, qo, ko, fast=True)
        diff = (res_fast.out.float() - ref.float())
        assert diff.norm() / ref.float().norm() < 4e-3
