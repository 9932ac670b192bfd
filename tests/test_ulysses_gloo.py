"""CPU, world_size 2 and 4 over gloo: the Ulysses head/sequence exchange is exact (layout logic only; the
attention in the middle is a per-head stand-in so no GPU is needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _per_head_op(x):
    """[1,S,h,D] -> [1,S,h,D]; depends on the whole sequence of each head (like attention does)."""
    w = torch.softmax(x.float().mean(-1, keepdim=True), dim=1)
    return (x.float() * w + x.float().flip(1).cumsum(1) * 1e-3).to(x.dtype)


def _worker(rank, world, degree, port, S, H, D):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from video_blade_b200.ulysses import UlyssesGroup
        ug = UlyssesGroup(world, rank, degree)
        gid, pr = rank // degree, rank % degree
        g = torch.Generator().manual_seed(100 + gid)            # one sequence per group
        full = torch.randn(1, S, H, D, generator=g)
        sl = slice(pr * (S // degree), (pr + 1) * (S // degree))
        x = full[:, sl].contiguous()
        xq, xk = ug.scatter_heads(x, x * 2)
        Hl = H // degree
        assert xq.shape == (1, S, Hl, D)
        assert torch.equal(xq, full[:, :, pr * Hl:(pr + 1) * Hl])
        assert torch.equal(xk, 2 * full[:, :, pr * Hl:(pr + 1) * Hl])
        y = ug.gather_heads(_per_head_op(xq))
        want = _per_head_op(full)[:, sl]
        assert y.shape == want.shape and torch.equal(y, want)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,degree", [(2, 2), (4, 2), (4, 4)])
def test_ulysses_roundtrip_gloo(world, degree):
    port = _free_port()
    mp.spawn(_worker, args=(world, degree, port, 24, 12, 8), nprocs=world, join=True)


def _worker_fused(rank, world, degree, port, S, H, D):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from video_blade_b200.ulysses import UlyssesGroup
        ug = UlyssesGroup(world, rank, degree)
        gid, pr = rank // degree, rank % degree
        g = torch.Generator().manual_seed(7 + gid)
        full = [torch.randn(1, S, H, D, generator=g) for _ in range(3)]
        sl = slice(pr * (S // degree), (pr + 1) * (S // degree))
        qv, kv, vv, vrow, keep = ug.scatter_heads_fused(*(x[:, sl].contiguous() for x in full))
        Hl = H // degree
        if degree == 1:
            # degree 1 must not touch the WORLD group (two CFG groups of one rank each): identity views of my sequence
            assert torch.equal(vrow, torch.arange(S, dtype=torch.int32))
            for view, ref in zip((qv, kv, vv), full):
                assert view.shape == (1, H, S, D) and torch.equal(view, ref.transpose(1, 2))
            return
        flat = keep.reshape(-1)
        for view, ref in zip((qv, kv, vv), full):
            assert view.shape == (1, Hl, S, D) and view.stride(2) == Hl * D and view.stride(1) == D
            # token s of the head shard is row vrow[s] of the strided view: the address arithmetic the prep kernel
            # does (base + vrow * stride_s + h * stride_h + d), emulated on the flat receive buffer
            off = view.storage_offset() + vrow.long()[None, :, None] * view.stride(2) \
                + torch.arange(Hl)[:, None, None] * view.stride(1) + torch.arange(D)[None, None, :]
            got = flat[off]                                                 # [Hl, S, D]
            want = ref[0, :, pr * Hl:(pr + 1) * Hl].transpose(0, 1)
            assert torch.equal(got, want)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,degree", [(2, 2), (4, 4), (2, 1), (4, 2)])
def test_ulysses_fused_scatter_layout_gloo(world, degree):
    port = _free_port()
    mp.spawn(_worker_fused, args=(world, degree, port, 24, 12, 8), nprocs=world, join=True)


# ------------------------------------------------------------------ CogVideoX scaffold, sequence-parallel blocks
class _DenseInner(torch.nn.Module):
    """Stand-in for the ASA module on CPU: plain attention, reading q/k/v out of the packed Ulysses receive buffer
    through the virtual-row table exactly like prep_block_kernel does on the GPU."""

    def forward(self, q, k, v, virtual_rows=None, rotary=None):
        if virtual_rows is not None:
            _, Hl, S, D = q.shape
            rows = virtual_rows.long()

            def gather(x):
                flat = x.as_strided((int(rows.max()) + 1, Hl, D), (Hl * D, D, 1), x.storage_offset())
                return flat[rows].permute(1, 0, 2)[None]                        # [1,Hl,S,D]
            q, k, v = gather(q), gather(k), gather(v)
        return torch.nn.functional.scaled_dot_product_attention(q, k, v)


def _worker_cog(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from video_blade_b200.dit import CogLikeDiT
        from video_blade_b200.modify_cogvideo import SageAttnCogVideoXAttnProcessor
        from video_blade_b200.ulysses import UlyssesGroup
        torch.manual_seed(0)
        model = CogLikeDiT(dim=64, heads=4, layers=2, text_dim=24, in_ch=4, patch=2, temb_dim=32).eval()
        inner = _DenseInner()
        for i, blk in enumerate(model.transformer_blocks):
            blk.attn1.inner_attention = inner
            blk.attn1.set_processor(SageAttnCogVideoXAttnProcessor(i, fuse_rope=False))
        g = torch.Generator().manual_seed(3)
        lat = torch.randn(2, 2, 4, 8, 8, generator=g)            # 2 x 4 x 4 = 32 video tokens
        txt = torch.randn(2, 6, 24, generator=g)                 # 6 text tokens -> 38 = 2 x 19
        t = torch.tensor([300.0, 700.0])
        with torch.no_grad():
            ref = model(lat, t, txt)
            model.set_sequence_parallel(UlyssesGroup(world, rank, world))
            got = model(lat, t, txt)
        assert got.shape == ref.shape
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), float((got - ref).abs().max())
    finally:
        dist.destroy_process_group()


def test_cog_scaffold_sequence_parallel_matches_single_rank_gloo():
    """[text ; video] sharded over 2 ranks: per-token text/video modulation, identity rotary rows for text, the
    fused q/k/v exchange and the final all-gather reproduce the unsharded forward."""
    mp.spawn(_worker_cog, args=(2, _free_port()), nprocs=2, join=True)
