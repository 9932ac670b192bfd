"""Processor / installer surface (modify_wan.py MW:75-168, modify_cogvideo.py MC:11-91).
CPU part: installers wire the objects exactly like the reference.  GPU part: a processor call through the
CUDA ASA layer equals the same processor with the oracle as `inner_attention`."""
import types

import pytest
import torch
import torch.nn as nn

from oracle import asa_oracle as O


@pytest.fixture(autouse=True)
def _inference_mode():
    """The CUDA path is forward-only and refuses inputs that require grad while autograd records
    (tests/test_host_logic.py::test_forward_only_guard): run the processors the way inference does."""
    with torch.no_grad():
        yield


class _Block(nn.Module):
    def __init__(self, dim, heads, qk_norm):
        super().__init__()
        from video_blade_b200.modify_wan import Attention
        self.attn1 = Attention(dim, heads, qk_norm=qk_norm)
        self.attn1.set_processor("stock")


def test_installers_wire_like_the_reference():
    from video_blade_b200 import modify_cogvideo as MC, modify_wan as MW
    model = types.SimpleNamespace(blocks=nn.ModuleList([_Block(64, 2, "rms_norm_across_heads") for _ in range(3)]))
    inner = MW.set_adaptive_block_sparse_attn_wanx(model, verbose=True)
    for b in model.blocks:
        assert b.attn1.inner_attention is inner                  # one shared module (MW:154-158)
        assert isinstance(b.attn1.get_processor(), MW.WanAttnProcessor2_0)
        assert b.attn1.origin_processor == "stock" and b.attn1.verbose is True
    model = types.SimpleNamespace(transformer_blocks=nn.ModuleList([_Block(64, 2, "layer_norm") for _ in range(2)]))
    inner = MC.set_block_sparse_attn_cogvideox(model)
    for i, b in enumerate(model.transformer_blocks):
        assert b.attn1.inner_attention is inner
        assert isinstance(b.attn1.get_processor(), MC.SageAttnCogVideoXAttnProcessor)
        assert b.attn1.get_processor().idx == i and b.attn1.origin_processor == "stock"


def test_wan_rope_fp32_matches_reference_fp64():
    from video_blade_b200.modify_wan import apply_rotary_emb
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 2, 50, 16, generator=g).bfloat16()
    ang = torch.rand(1, 1, 50, 8, generator=g, dtype=torch.float64) * 6.28
    freqs = torch.polar(torch.ones_like(ang), ang)
    ref = O.apply_rotary_emb_wan(x, freqs)
    got = apply_rotary_emb(x, freqs.to(torch.complex64))
    assert (got.float() - ref.float()).abs().max() <= 2 ** -6      # <= 1 bf16 ulp at |x| < 4


@pytest.mark.gpu
def test_wan_processor_end_to_end():
    from video_blade_b200 import modify_wan as MW, wanx_blocksparseattn as W
    grid = (26, 15, 4)
    S, H, D = grid[0] * grid[1] * grid[2], 2, 128
    W.width, W.height, W.depth, W.max_retain_ratio = *grid, 0.4
    try:
        torch.manual_seed(0)
        model = types.SimpleNamespace(blocks=nn.ModuleList([_Block(H * D, H, "rms_norm_across_heads")]))
        model.blocks.to("cuda", torch.bfloat16)
        inner = MW.set_adaptive_block_sparse_attn_wanx(model)
        inner.print_every = 0
        attn = model.blocks[0].attn1
        x = torch.randn(1, S, H * D, device="cuda", dtype=torch.bfloat16)
        ang = torch.rand(1, 1, S, D // 2, device="cuda") * 6.28
        freqs = torch.polar(torch.ones_like(ang), ang)
        out = attn(x, rotary_emb=freqs)                                  # rotary embedding fused into the gather kernel
        assert out.shape == x.shape and out.dtype == x.dtype and torch.isfinite(out.float()).all()
        attn.set_processor(MW.WanAttnProcessor2_0(fuse_rope=False))     # same thing with the torch rotary path
        out_t = attn(x, rotary_emb=freqs)
        dd = out.float() - out_t.float()
        assert float(dd.norm() / out_t.float().norm()) <= 1e-2, float(dd.norm() / out_t.float().norm())

        # same processor, oracle as inner attention (fed the kernel's own fp32 scores)
        captured = {}

        class OracleInner(nn.Module):
            def forward(self, q, k, v):
                eng = W._engine()
                _, dbg = eng.forward(q, k, v, return_debug=True)
                cfg = O.ASAConfig.wan(width=grid[0], height=grid[1], depth=grid[2], max_retain_ratio=0.4)
                res = O.asa_forward(q.cpu(), k.cpu(), v.cpu(), cfg, scores=dbg["scores"].cpu())
                captured["mask_equal"] = torch.equal(res.mask, dbg["mask"].cpu())
                return res.out.to(q.device)
        attn.inner_attention = OracleInner()
        ref = attn(x, rotary_emb=freqs)
        assert captured["mask_equal"]
        d = out.float() - ref.float()
        assert float(d.norm() / ref.float().norm()) <= 1e-2 and float(d.abs().max()) <= 2e-2
    finally:
        W.width, W.height, W.depth, W.max_retain_ratio = 52, 30, 21, 0.17


@pytest.mark.gpu
def test_cog_processor_end_to_end():
    from video_blade_b200 import cogvideo_blocksparseattn as C, modify_cogvideo as MC
    grid, T = (15, 10, 6), 40
    Sv, H, D = grid[0] * grid[1] * grid[2], 3, 64
    C.width, C.height, C.depth, C.text_length, C.max_retain_ratio = *grid, T, 0.3
    try:
        torch.manual_seed(1)
        model = types.SimpleNamespace(transformer_blocks=nn.ModuleList([_Block(H * D, H, "layer_norm")]))
        model.transformer_blocks.to("cuda", torch.bfloat16)
        inner = MC.set_block_sparse_attn_cogvideox(model)
        inner.print_every = 0
        attn = model.transformer_blocks[0].attn1
        x = torch.randn(1, Sv, H * D, device="cuda", dtype=torch.bfloat16)
        txt = torch.randn(1, T, H * D, device="cuda", dtype=torch.bfloat16)
        ang = torch.rand(Sv, D // 2, device="cuda") * 6.28
        cos, sin = ang.cos().repeat_interleave(2, -1), ang.sin().repeat_interleave(2, -1)
        hs, ehs = attn(x, encoder_hidden_states=txt, image_rotary_emb=(cos, sin))
        assert hs.shape == x.shape and ehs.shape == txt.shape
        assert torch.isfinite(hs.float()).all() and torch.isfinite(ehs.float()).all()
        attn.set_processor(MC.SageAttnCogVideoXAttnProcessor(0, fuse_rope=False))
        hs_t, ehs_t = attn(x, encoder_hidden_states=txt, image_rotary_emb=(cos, sin))
        for a_, b_ in ((hs, hs_t), (ehs, ehs_t)):
            dd = a_.float() - b_.float()
            assert float(dd.norm() / b_.float().norm()) <= 1e-2
    finally:
        C.width, C.height, C.depth, C.text_length, C.max_retain_ratio = 45, 30, 13, 226, 0.1


@pytest.mark.gpu
def test_wan_processor_fused_qk_norm_matches_torch_norm():
    """RMSNorm over all heads' channels (MW:99-102) inside the gather kernel == the module's own forward followed by
    the unfused path, up to bf16 rounding of the normalised q/k."""
    import torch
    from video_blade_b200 import wanx_blocksparseattn as W
    from video_blade_b200.modify_wan import Attention, WanAttnProcessor2_0
    from video_blade_b200.dit import rope_freqs
    torch.manual_seed(0)
    old = (W.width, W.height, W.depth)
    W.width, W.height, W.depth = 26, 15, 8
    try:
        S, dim, heads = 26 * 15 * 8, 512, 4
        attn = Attention(dim, heads, qk_norm="rms_norm_across_heads").cuda().to(torch.bfloat16)
        with torch.no_grad():
            attn.norm_q.weight.copy_(1 + 0.3 * torch.randn(dim))
            attn.norm_k.weight.copy_(1 + 0.3 * torch.randn(dim))
        inner = W.AdaptiveBlockSparseAttnTrain()
        inner.print_every = 0
        attn.inner_attention = inner
        x = torch.randn(2, S, dim, device="cuda", dtype=torch.bfloat16)
        rope = rope_freqs(8, 15, 26, dim // heads, device="cuda")
        outs = []
        for fuse in (False, True):
            attn.set_processor(WanAttnProcessor2_0(fuse_norm=fuse))
            with torch.no_grad():
                outs.append(attn(x, rotary_emb=rope).float())
        d = outs[1] - outs[0]
        rel = float(d.norm() / outs[0].norm())
        assert rel < 1e-2, rel
    finally:
        W.width, W.height, W.depth = old


@pytest.mark.gpu
def test_cog_processor_fused_layernorm_matches_torch_norm():
    """Per-head LayerNorm of q/k (MC:54-57) inside the gather kernel == the modules' own forward followed by the
    unfused-norm path (all blocks retained, so only the normalisation arithmetic differs)."""
    import torch
    from video_blade_b200 import cogvideo_blocksparseattn as Cg
    from video_blade_b200.modify_wan import Attention
    from video_blade_b200.modify_cogvideo import SageAttnCogVideoXAttnProcessor
    from video_blade_b200.dit import rope_cos_sin
    torch.manual_seed(0)
    old = (Cg.width, Cg.height, Cg.depth, Cg.text_length, Cg.max_retain_ratio, Cg.min_retain_ratio)
    Cg.width, Cg.height, Cg.depth, Cg.text_length = 15, 10, 6, 40
    Cg.max_retain_ratio = Cg.min_retain_ratio = 1.0
    try:
        Sv, T, dim, heads = 15 * 10 * 6, 40, 192, 3
        attn = Attention(dim, heads, qk_norm="layer_norm").cuda().to(torch.bfloat16)
        with torch.no_grad():
            for n in (attn.norm_q, attn.norm_k):
                n.weight.copy_(1 + 0.3 * torch.randn(dim // heads))
                n.bias.copy_(0.2 * torch.randn(dim // heads))
        inner = Cg.AdaptiveBlockSparseAttnTrain()
        inner.print_every = 0
        attn.inner_attention = inner
        x = torch.randn(2, Sv, dim, device="cuda", dtype=torch.bfloat16)
        txt = torch.randn(2, T, dim, device="cuda", dtype=torch.bfloat16)
        rope = rope_cos_sin(6, 10, 15, dim // heads, device="cuda")
        outs = []
        for fuse in (False, True):
            attn.set_processor(SageAttnCogVideoXAttnProcessor(0, fuse_norm=fuse))
            with torch.no_grad():
                hv, ht = attn(x, encoder_hidden_states=txt, image_rotary_emb=rope)
            outs.append(torch.cat([ht, hv], 1).float())
        d = outs[1] - outs[0]
        rel = float(d.norm() / outs[0].norm())
        assert rel < 1e-2, rel
    finally:
        Cg.width, Cg.height, Cg.depth, Cg.text_length, Cg.max_retain_ratio, Cg.min_retain_ratio = old


def test_norm_fusion_is_only_claimed_for_known_modules():
    """The fused q/k norm reproduces a specific module's rounding; unknown normalisation modules stay on torch."""
    import torch
    from video_blade_b200.modify_wan import RMSNorm, _rms_kind
    assert _rms_kind(RMSNorm(8)) == 1
    assert _rms_kind(torch.nn.LayerNorm(8)) == 0
    if hasattr(torch.nn, "RMSNorm"):
        assert _rms_kind(torch.nn.RMSNorm(8)) == 0

    class Fake(torch.nn.Module):
        pass
    Fake.__name__ = "RMSNorm"
    Fake.__module__ = "diffusers.models.normalization"
    assert _rms_kind(Fake()) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("estimator", ["sampled_max", "meanpool"])
def test_wan_processor_installer_defaults_with_each_estimator(estimator):
    """The installer's default processor (fused rope + fused RMSNorm) must work with the reference's estimator too:
    the one-call layer applies norm and rotation in the gather kernel for every estimator (the staged
    sampled-max path used to raise on the fused norm)."""
    from video_blade_b200 import modify_wan as MW, wanx_blocksparseattn as W
    grid = (26, 15, 4)
    S, H, D = grid[0] * grid[1] * grid[2], 2, 128
    W.width, W.height, W.depth, W.max_retain_ratio, W.estimator = *grid, 0.4, estimator
    try:
        torch.manual_seed(0)
        model = types.SimpleNamespace(blocks=nn.ModuleList([_Block(H * D, H, "rms_norm_across_heads")]))
        model.blocks.to("cuda", torch.bfloat16)
        inner = MW.set_adaptive_block_sparse_attn_wanx(model)
        inner.print_every = 0
        attn = model.blocks[0].attn1
        x = torch.randn(1, S, H * D, device="cuda", dtype=torch.bfloat16)
        ang = torch.rand(1, 1, S, D // 2, device="cuda") * 6.28
        freqs = torch.polar(torch.ones_like(ang), ang)
        torch.manual_seed(5)                                              # the module draws the sample offsets (W:49-51)
        out = attn(x, rotary_emb=freqs)
        attn.set_processor(MW.WanAttnProcessor2_0(fuse_rope=False, fuse_norm=False))
        torch.manual_seed(5)
        out_t = attn(x, rotary_emb=freqs)
        assert torch.isfinite(out.float()).all()
        dd = out.float() - out_t.float()
        # fused vs torch norm/rope differ by bf16 roundings of q/k; with the sampled estimator that can move a block in
        # or out of a row's selection, so the bound is the loose one
        assert float(dd.norm() / out_t.float().norm()) <= (3e-2 if estimator == "sampled_max" else 1e-2)
        assert 0.0 < inner.average_sparsity() < 1.0                       # device counter fed by the selection kernel
    finally:
        W.width, W.height, W.depth, W.max_retain_ratio, W.estimator = 52, 30, 21, 0.17, "meanpool"


@pytest.mark.gpu
def test_scaffold_fused_glue_matches_torch_expressions():
    """The benchmark scaffold's fused token-wise kernels (scaffold_ops, not the hot path) against the torch expressions
    they replace: identical up to one bf16 rounding."""
    from video_blade_b200 import scaffold_ops as ops
    torch.manual_seed(0)
    for C, S in ((1536, 517), (3072, 130), (256, 33)):
        x = torch.randn(2, S, C, device="cuda").bfloat16()
        y = torch.randn(2, S, C, device="cuda").bfloat16()
        sc, sh, g = (torch.randn(2, 1, C, device="cuda") * 0.3 for _ in range(3))
        want = (torch.nn.functional.layer_norm(x.float(), (C,), eps=1e-6) * (1 + sc) + sh)
        got = ops.ln_modulate(x, sc, sh, 1e-6)
        def one_ulp(a, b):                       # |a - b| within one bf16 ulp of the value
            return bool(((a.float() - b.float()).abs() <= b.float().abs() * 2.0 ** -7 + 1e-6).all())
        assert got.dtype == x.dtype and one_ulp(got, want)
        assert float((got.float() - want.bfloat16().float()).abs().gt(0).float().mean()) < 0.02
        lw, lb = (1 + 0.2 * torch.randn(C, device="cuda")).bfloat16(), (0.1 * torch.randn(C, device="cuda")).bfloat16()
        want = torch.nn.functional.layer_norm(x.float(), (C,), lw.float(), lb.float(), eps=1e-5) * (1 + sc) + sh
        assert one_ulp(ops.ln_modulate(x, sc, sh, 1e-5, lw, lb), want)      # affine variant (CogVideoX LayerNormZero)
        want = (x.float() + y.float() * g).bfloat16()
        got = ops.gated_residual(x, y, g)
        assert float((got.float() - want.float()).abs().gt(0).float().mean()) < 0.01
        assert one_ulp(got, want)
        w = (1 + 0.2 * torch.randn(C, device="cuda")).bfloat16()
        v = x.float()
        want = (v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()).bfloat16()
        got = ops.rmsnorm(x, w, 1e-6)
        assert float((got.float() - want.float()).abs().gt(0).float().mean()) < 0.02
        assert one_ulp(got, want)
    lin = torch.nn.Linear(512, 1024).cuda().bfloat16()
    xx = torch.randn(3, 77, 512, device="cuda").bfloat16()
    want = torch.nn.functional.gelu(lin(xx), approximate="tanh")
    got = ops.linear_gelu_tanh(xx, lin)
    assert got.shape == want.shape and float((got.float() - want.float()).abs().max()) <= 3e-2
