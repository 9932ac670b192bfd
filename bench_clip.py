#!/usr/bin/env python
"""End-to-end 8-step clip benchmark (BASELINE.json config 3): random-init Wan2.1-T2V-1.3B-shaped DiT (30 blocks),
synthetic latents [B,16,21,60,104] + prompt embeddings [B,512,4096], CFG, `generate_new` sampler, ASA installed
through the reference-facing installer.  1 GPU: CFG as a batch of two.  N GPUs (torchrun): 2 CFG groups x
Ulysses N/2 (Wan has 12 heads -> N in {2,4,8}).

    python bench_clip.py [--steps 8] [--layers 30] [--reps 2]
    python -m torch.distributed.run --nproc-per-node N ... bench_clip.py --gpus N

Prints one JSON line: clip seconds (max over ranks, CUDA events), and the attention share.  `run_clip` is also what
bench.py calls to put BASELINE's second metric ("8-step clip s at 1/2/4/8 GPU") on the driver-run line."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run_clip(world, rank, dev, steps=8, layers=30, reps=2, model="wan", guidance=5.0, dense=False, retain=None,
             data_plane="auto", sampler="generate_new", hoist=False):
    """One process per GPU; the process group (NCCL) must exist when world > 1.  Returns the result dict on every
    rank (value = max over ranks of the best repetition, seconds)."""
    import torch.distributed as dist
    from video_blade_b200.dit import WanLikeDiT, generate_new, make_velocity_fn
    from video_blade_b200.modify_wan import set_adaptive_block_sparse_attn_wanx
    from video_blade_b200.ulysses import UlyssesGroup

    P = 1 if world == 1 else world // 2
    group = UlyssesGroup(world, rank, P) if world > 1 else None
    branch = 0 if world == 1 else rank // P
    torch.manual_seed(0)

    class Dense(torch.nn.Module):
        def forward(self, q, k, v, **kw):
            return torch.nn.functional.scaled_dot_product_attention(q, k, v)
    g = torch.Generator(device="cpu").manual_seed(1)
    if model == "cog":
        from video_blade_b200 import cogvideo_blocksparseattn as Cg
        from video_blade_b200.dit import CogLikeDiT
        from video_blade_b200.modify_cogvideo import set_block_sparse_attn_cogvideox
        if retain is not None:
            Cg.max_retain_ratio = Cg.min_retain_ratio = retain
        layers = 42 if layers == 30 else layers
        with torch.device(dev):
            net = CogLikeDiT(layers=layers).to(torch.bfloat16).eval()
        blocks = net.transformer_blocks
        inner = set_block_sparse_attn_cogvideox(net)
        net.set_sequence_parallel(group, data_plane=data_plane)
        noise = torch.randn(1, 13, 16, 60, 90, generator=g).to(dev, torch.bfloat16)
        prompt = torch.randn(1, 226, 4096, generator=g).to(dev, torch.bfloat16)
        negative = torch.randn(1, 226, 4096, generator=g).to(dev, torch.bfloat16)
        wname = f"CogVideoX-5B 49x480x720, {layers} DiT blocks, {steps} steps, CFG batch 2" + \
                (f", min=max retain {retain}" if retain is not None else "")
    else:
        net = WanLikeDiT(layers=layers).to(dev, torch.bfloat16).eval()
        blocks = net.blocks
        inner = set_adaptive_block_sparse_attn_wanx(net)
        net.set_sequence_parallel(group, data_plane=data_plane)
        if hoist:
            net.set_hoisted_permutation(True)
        noise = torch.randn(1, 16, 21, 60, 104, generator=g).to(dev, torch.bfloat16)
        prompt = torch.randn(1, 512, 4096, generator=g).to(dev, torch.bfloat16)
        negative = torch.randn(1, 512, 4096, generator=g).to(dev, torch.bfloat16)
        wname = f"Wan2.1-T2V-1.3B 81x480x832, {layers} DiT blocks, {steps} steps, CFG batch 2"
    inner.print_every = 0
    if dense:
        for blk in blocks:
            blk.attn1.inner_attention = Dense()
    cfg_ranks = None if world == 1 else (branch, 0, P)
    vel = make_velocity_fn(net, prompt, negative, guidance, cfg_ranks)

    if sampler == "unipc":                                       # WanPipeline + UniPC-flow (inference.py:48-52)
        from video_blade_b200.samplers import sample_unipc_flow
        generate_new = lambda f, x, steps: sample_unipc_flow(f, x, steps=steps)          # noqa: E731
    elif sampler == "dpm":                                       # CogVideoXPipeline + DPM "trailing" (inference.py:64-66)
        from video_blade_b200.samplers import sample_cogvideox_dpm
        gen_n = torch.Generator(device=dev).manual_seed(3)
        generate_new = lambda f, x, steps: sample_cogvideox_dpm(f, x, steps=steps, generator=gen_n)   # noqa: E731
    with torch.no_grad():
        generate_new(vel, noise, steps=1)                        # warm-up (cuBLAS plans, workspaces, NCCL)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            out = generate_new(vel, noise, steps=steps)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 1e3)
    t = torch.tensor([min(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    plane = getattr(net, "data_plane_in_use", "none")
    res = {"metric": f"8-step clip seconds ({'CogVideoX-5B' if model == 'cog' else 'Wan2.1-T2V-1.3B'} shape, random init, CFG, synthetic inputs)",
           "value": float(t.item()), "unit": "s", "higher_is_better": False, "n_gpus": world,
           "config": {"workload": wname, "layers": layers, "steps": steps, "attention": "dense SDPA" if dense else "ASA",
                      "parallelism": "single (CFG batch 2)" if world == 1 else f"cfg2xulysses{P}", "data_plane": plane,
                      "sampler": sampler, "gilbert_permutation": "hoisted to model level" if hoist else "per layer (reference)"},
           "finite": bool(torch.isfinite(out.float()).all()), "all_reps_s": times,
           "avg_sparsity": None if dense else inner.average_sparsity()}
    del net, vel, out
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--layers", type=int, default=30)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--guidance", type=float, default=5.0)
    ap.add_argument("--dense", action="store_true", help="replace ASA by dense SDPA (context number)")
    ap.add_argument("--model", default="wan", choices=["wan", "cog"], help="wan = config 3, cog = config 5")
    ap.add_argument("--retain", type=float, default=None, help="cog density sweep: min = max retain ratio")
    ap.add_argument("--data-plane", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--hoist", action="store_true", help="Wan: permute tokens into curve order once per forward (SURVEY 7.3)")
    ap.add_argument("--sampler", default="generate_new", choices=["generate_new", "unipc", "dpm"],
                    help="generate_new = the trainer's K-step rollout (TW:1402-1443); unipc / dpm = the inference scripts' schedulers")
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == a.gpus
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run_clip(world, rank, dev, steps=a.steps, layers=a.layers, reps=a.reps, model=a.model, guidance=a.guidance,
                   dense=a.dense, retain=a.retain, data_plane=a.data_plane, sampler=a.sampler, hoist=a.hoist)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
