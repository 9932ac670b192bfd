"""CPU: the multi-level pooled-mask oracle (forward) against the reference's own code --
its `transfer_attn_to_mask` (N:154-207) and its Triton kernel K9 run under the interpreter
(oracle/make_golden_multilevel.py)."""
import numpy as np
import pytest
import torch

from conftest import load_npz
from oracle import multilevel as M


@pytest.fixture(scope="module")
def gold():
    return load_npz("multilevel.npz")[0]


@pytest.mark.parametrize("name", ["m_small", "m_cog"])
def test_multilevel_mask_bit_exact(gold, name):
    attn = torch.from_numpy(gold[f"{name}_attn"])
    want = torch.from_numpy(gold[f"{name}_mask"])
    got = M.multilevel_mask(attn)
    assert got.dtype == torch.int32 and torch.equal(got, want)
    assert set(np.unique(got.numpy())) <= {0, 1, 2, 4, 8}
    assert bool((got[..., -2:, :] == 1).all()) and bool((got[..., :, -2:] == 1).all())
    # the ratios the reference module actually runs with (N:13-19): 5 % full, 10 % 2x, 10 % 4x, 25 % 8x, rest skipped
    ratios = {int(lv): (float(a), float(b)) for lv, a, b in gold["module_ratios"]}
    assert ratios[0] == (0.5, 1.0)
    assert torch.equal(M.multilevel_mask(attn, ratios), torch.from_numpy(gold[f"{name}_mask_module_ratios"]))


@pytest.mark.parametrize("name", ["a_levels", "a_ragged", "a_d128"])
def test_multilevel_attention_matches_reference_kernel(gold, name):
    q, k, v, mask, want = (torch.from_numpy(gold[f"{name}_{x}"]) for x in ("q", "k", "v", "mask", "o"))
    got = M.multilevel_attention(q, k, v, mask)
    d = (got - want).abs()
    assert float(d.max()) < 2e-5, float(d.max())          # fp32 on both sides: only summation order differs


def test_levels_zero_and_one_reduce_to_block_masked_attention():
    """With levels in {0, 1} and a length that is a multiple of 128 the multi-level kernel is plain block-masked
    attention -- the same definitional form the main oracle uses for the external kernel."""
    from oracle import asa_oracle as O
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(1, 2, 512, 64, generator=g) for _ in range(3))
    mask = torch.rand(1, 2, 4, 4, generator=g) < 0.5
    mask |= torch.eye(4, dtype=torch.bool)
    got = M.multilevel_attention(q, k, v, mask.to(torch.int32))
    want, _ = O.dense_masked_attention(q, k, v, mask)
    assert float((got - want).abs().max()) < 2e-5


def test_pooled_level_equals_attention_over_pooled_keys_with_log_bias():
    """One key block at level 4 == softmax over its 32 four-token means with +log 4 (K9:183-229)."""
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(1, 1, 128, 64, generator=g) for _ in range(3))
    got = M.multilevel_attention(q, k, v, torch.full((1, 1, 1, 1), 4, dtype=torch.int32))
    k4 = k.view(1, 1, 32, 4, 64).mean(3)
    v4 = v.view(1, 1, 32, 4, 64).mean(3)
    want = torch.softmax(q @ k4.transpose(-1, -2) / 8.0, -1) @ v4     # a constant bias cancels in one softmax
    assert float((got - want).abs().max()) < 2e-5


def test_multilevel_layer_matches_reference_module(gold):
    """N's AdaptiveBlockSparseAttnTrain.forward end to end (rearrangement, sampled-max estimator, multi-level mask,
    multi-level attention, inverse rearrangement), its random sampling draws made explicit."""
    q, k, v, want = (torch.from_numpy(gold[f"layer_{x}"]) for x in ("q", "k", "v", "o"))
    w, h, d, text = (int(x) for x in gold["layer_grid_text"])
    ratios = {int(lv): (float(a), float(b)) for lv, a, b in gold["layer_ratios"]}
    offs = [torch.topk(torch.from_numpy(gold[f"layer_rand_{t}"]), 32, dim=3).indices[:, :, 0, :] for t in ("q", "k")]
    got, mask, _ = M.multilevel_forward(q, k, v, (w, h, d), text, ratios, offs[0], offs[1])
    assert set(np.unique(mask.numpy())) <= {0, 1, 2, 4, 8} and int((mask == 0).sum()) > 0
    assert float((got - want).abs().max()) < 5e-5


def test_multilevel_backward_is_the_gradient_of_the_oracle(gold):
    """The reference's hand-written backward kernels (K9:695-1237) compute the plain gradient of the forward,
    including the path through the pooled K/V copies: autograd through the oracle reproduces dq, dk, dv."""
    q, k, v = (torch.from_numpy(gold[f"bwd_{x}"]).clone().requires_grad_(True) for x in ("q", "k", "v"))
    mask, do = torch.from_numpy(gold["bwd_mask"]), torch.from_numpy(gold["bwd_do"])
    o = M.multilevel_attention(q, k, v, mask)
    assert float((o.detach() - torch.from_numpy(gold["bwd_o"])).abs().max()) < 2e-5
    o.backward(do)
    for name, grad in (("dq", q.grad), ("dk", k.grad), ("dv", v.grad)):
        want = torch.from_numpy(gold[f"bwd_{name}"])
        assert float((grad - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max())), name
