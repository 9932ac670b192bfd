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


def _replay_schedule(B, H, nq, npt, G, cnt, dynamic=True, half_tiles=True):
    import ctypes as C
    import numpy as np
    from video_blade_b200 import _lib
    lib = _lib.load()
    cnt = np.ascontiguousarray(cnt, dtype=np.int32)
    cap = 4 * B * H * ((nq + 1) // 2) + 8
    items = np.zeros((cap, 12), dtype=np.int32)
    n = C.c_int32(0)
    _lib.check(lib.blade_debug_attn_schedule(B, H, nq, npt, G, cnt.ctypes.data, int(dynamic), int(half_tiles),
                                             items.ctypes.data, cap, C.byref(n)))
    return items[:n.value]


@pytest.mark.parametrize("B,H,nq,npt,G", [(1, 12, 256, 9, 148), (1, 3, 256, 9, 148), (1, 48, 139, 10, 148), (2, 5, 7, 3, 148),
                                          (1, 1, 1, 1, 148), (1, 2, 33, 0, 148), (1, 1, 16, 1, 148), (3, 7, 64, 2, 132),
                                          (1, 40, 64, 9, 148), (1, 1, 300, 2, 8)])
@pytest.mark.parametrize("dynamic,half_tiles", [(True, True), (True, False), (False, False)])
def test_attention_schedule_covers_every_tile_exactly_once(B, H, nq, npt, G, dynamic, half_tiles):
    """Host replay of the persistent attention kernel's item decode (the same classify / make_item code the device roles
    run, `blade_debug_attn_schedule`): for every (batch*head, query tile) the pooled tiles are run exactly once and the
    slices [off, off + n) of the block list handed to streams / CTAs partition [0, count) -- for tile pairs, solo tiles
    and half tiles, odd tile counts, short lists and fewer pairs than SMs."""
    import numpy as np
    rng = np.random.default_rng(B * 1000 + H * 10 + nq)
    cnt = rng.integers(1, 45, size=(B, H, nq), dtype=np.int32)
    cnt[..., : max(1, nq // 7)] = rng.integers(1, 5, size=(B, H, max(1, nq // 7)))     # lists too short to split
    items = _replay_schedule(B, H, nq, npt, G, cnt, dynamic, half_tiles)
    pooled = np.zeros((B * H, nq), dtype=np.int64)
    cover = [[[] for _ in range(nq)] for _ in range(B * H)]
    halves = {}
    for row in items:
        item, bh, merge, split = int(row[0]), int(row[1]), int(row[10]), int(row[11])
        assert 0 <= bh < B * H
        streams = [tuple(int(x) for x in row[2 + 4 * t: 6 + 4 * t]) for t in range(2)]     # (qb, pt, off, ns)
        for t, (qb, pt, off, ns) in enumerate(streams):
            assert 0 <= qb <= nq and pt in (0, npt) and off >= 0 and ns >= 0
            if qb == nq:
                assert pt == 0 and ns == 0                                             # no tile: no work
                continue
            pooled[bh, qb] += pt
            if ns:
                cover[bh][qb].append((off, ns))
        if merge:                                   # two streams of ONE tile: stream 1 continues stream 0's slice
            assert streams[0][0] == streams[1][0] and streams[1][2] == streams[0][2] + streams[0][3] and streams[1][3] > 0
        if split:
            slot, half = (split - 1) // 2, (split - 1) % 2
            assert merge and streams[0][3] + streams[1][3] >= 2                        # both halves work on both streams
            assert (slot, half) not in halves and 0 <= slot < min(G, 496)
            halves[(slot, half)] = (bh, streams[0][0])
    for (slot, half), tile in halves.items():
        assert halves.get((slot, 1 - half)) == tile, "a half tile without its partner"
    if not (dynamic and half_tiles):
        assert not halves
    assert (pooled == npt).all()
    for bh in range(B * H):
        b, h = divmod(bh, H)
        for qb in range(nq):
            segs = sorted(cover[bh][qb])
            pos = 0
            for off, ns in segs:
                assert off == pos, (bh, qb, segs)
                pos += ns
            assert pos == int(cnt[b, h, qb]), (bh, qb, segs, int(cnt[b, h, qb]))
