"""Host-side behaviour that needs no GPU: the forward-only guard, knob plumbing of the two mirrors, the sampler
restatements' bookkeeping."""
import pytest
import torch


def test_forward_only_guard():
    """the C ABI has no backward; a caller whose q/k/v require grad must get an error, not a silent
    gradient-free output.  The guard fires before anything touches CUDA, so it is testable here."""
    from video_blade_b200 import cogvideo_blocksparseattn as Cg, wanx_blocksparseattn as W
    for M in (W, Cg):
        layer = M.AdaptiveBlockSparseAttnTrain()
        q = torch.randn(1, 1, 8, 128, dtype=torch.bfloat16, requires_grad=True)
        with pytest.raises(RuntimeError, match="forward-only"):
            layer(q, q, q)
        with pytest.raises(RuntimeError, match="forward-only"):
            M.block_sparse_attn(q, q, q, torch.ones(1, 1, 1, 1, dtype=torch.bool))
        with pytest.raises(RuntimeError, match="forward-only"):
            M.adaptive_block_sparse_attn(q, q, q)
        with torch.no_grad():                                   # no grad recording -> the guard lets the call through
            with pytest.raises(RuntimeError, match="CUDA tensors only"):   # ... to the "no CPU fallback" error
                layer(q, q, q)


def test_mirrors_share_one_implementation_but_keep_their_knobs():
    from video_blade_b200 import cogvideo_blocksparseattn as Cg, wanx_blocksparseattn as W
    assert (W.max_retain_ratio, W.sample_gap, W.text_length, W.width, W.height, W.depth) == (0.17, 30, 0, 52, 30, 21)
    assert (Cg.max_retain_ratio, Cg.sample_gap, Cg.text_length, Cg.width, Cg.height, Cg.depth) == (0.1, 15, 226, 45, 30, 13)
    assert W.AdaptiveBlockSparseAttnTrain is not Cg.AdaptiveBlockSparseAttnTrain
    assert W.AdaptiveBlockSparseAttnTrain.__module__.endswith("wanx_blocksparseattn")
    kw, kc = W._knobs(), Cg._knobs()
    assert kw.flavor == "wan" and kc.flavor == "cog" and kc.c_config(139).force_last == 2 and kw.c_config(256).force_last == 0
    W.max_retain_ratio = 0.3                                    # knobs are read at call time, like the reference globals
    try:
        assert W._knobs().max_retain_ratio == 0.3 and Cg._knobs().max_retain_ratio == 0.1
    finally:
        W.max_retain_ratio = 0.17
    # cog rearranger moves the text rows to the tail and back (C:141-161)
    r = Cg.GilbertRearranger(4, 3, 2, text_length=5)
    x = torch.arange(29.0).view(1, 1, 29, 1)
    a, _, _ = r.rearrange(x, x, x)
    assert a[0, 0, -5:, 0].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0]
    assert torch.equal(r.reversed_rearrange(a), x)


def test_select_rounding_knob_reaches_the_c_config():
    from video_blade_b200.asa import AsaKnobs
    assert AsaKnobs.wan().c_config(256).select_rounding == 0
    assert AsaKnobs.wan(select_rounding="bf16").c_config(256).select_rounding == 1
    assert AsaKnobs.cog(select_rounding="f16").c_config(139).select_rounding == 2
