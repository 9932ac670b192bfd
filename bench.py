#!/usr/bin/env python
"""bench.py -- the driver contract for the ASA hot path (BASELINE.json: "ASA ms/layer + sparse-eff TFLOP/s
@Wan 32760 tok").

One "step" = one Adaptive-Sparse-Attention layer call (`AdaptiveBlockSparseAttnTrain.forward`, W:383-408)
on synthetic bf16 q,k,v of BASELINE config 2: Wan2.1-T2V-1.3B, 81x480x832 -> [B,12,32760,128], Gilbert
rearrangement on, block 128, retain 5-17 %, energy 0.95, pooled branch gap 30.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl blade|reference] [--inputs gaussian|structured]

N = 1 : the workload is exactly config 2 (B = 1).
N > 1 : torchrun, one rank per GPU over NCCL.  Headline (`value`): weak scaling by CFG / prompt batch split --
        global batch B = N sequences, one per rank, no data-path collective (ASA is independent per batch
        element and head).  In the same run the Ulysses configuration of config 3 is measured and reported under
        "ulysses": B = 2 (the CFG pair) as 2 groups x Ulysses degree N/2 (Wan: 12 heads, N=8 -> 2 x 4), a
        head/sequence all-to-all (NCCL) before and after the attention call, exchange time broken out.
        Everything is timed on the device, max over ranks.
`value` = algorithmic sparse-attention FLOPs of all layers processed per second,
inputs resident in HBM.  `e2e` = the same metric through the host-facing call: pinned host q,k,v -> H2D ->
layer -> D2H of the output, every step.

--impl reference times the reference's own CPU path (dense-masked PyTorch substitute for the absent CUDA
library; the oracle port, since the Python reference tree does not travel to the GPU box) on the host cores,
on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "Wan2.1-T2V-1.3B ASA layer 81x480x832 (32760 tok, 12 heads, d=128), block 128, retain 0.05-0.17, gap 30"
WORKLOAD_COG = ("CogVideoX-5B ASA layer 49x480x720 (17550 video + 226 text tok, 48 heads, d=64), block 128, "
                "retain 0.05-0.10, gap 15, last two block rows/cols dense")


def workload(args):
    """(knobs, H, D, name) of the requested BASELINE config: wan = config 2 (default), cog = config 4."""
    from video_blade_b200.asa import AsaKnobs
    if args.workload == "cog":
        kn, H, D, name = AsaKnobs.cog(), 48, 64, WORKLOAD_COG
    else:
        kn, H, D, name = AsaKnobs.wan(), 12, 128, WORKLOAD
    if args.retain is not None:                      # config 5 sweep: pin the density with min = max
        kn.max_retain_ratio = kn.min_retain_ratio = args.retain
        name += f", min=max retain {args.retain}"
    return kn, H, D, name


# --------------------------------------------------------------------------------------------- helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def traffic_bytes():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)["asa_attn_kernel<128,bf16>"]["dram_bytes_per_launch"]
    except Exception:
        return None


def make_inputs(B, H, S, D, kind, seed, grid, text_length=0):
    from video_blade_b200.synth import synth_qkv     # synthetic-input recipe shared with the tests (not timed)
    q, k, v = synth_qkv(B, H, S, D, seed=seed, structured=2.0 if kind == "structured" else 0.0, grid=grid,
                        text_length=text_length)
    # the reference hands the module transposed views of [B,S,H,D] memory (modify_wan.py:104-106)
    return tuple(x.transpose(1, 2).contiguous() for x in (q, k, v))     # [B,S,H,D] host tensors


def algorithmic_flops(cnt, mask_last_col, S, D, n_pool, block=128):
    """4*D*sum rows_i*(sum selected cols_j + n_pool), ragged tails at true size."""
    from video_blade_b200.synth import attention_flops
    nb = cnt.shape[-1]
    cols = cnt.float() * block
    tail = S - (nb - 1) * block
    cols = cols - mask_last_col.float() * (block - tail)
    return attention_flops(cols.cpu(), S, D, block, n_pool)


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_sample(q, k, v, kn, heads=1, threads=None):
    """The reference's CPU path (oracle port: reference module code + dense-masked substitute) on `heads`
    heads of the workload.  Returns (seconds, algorithmic FLOPs of the sample, threads)."""
    from oracle import asa_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = O.ASAConfig(flavor=kn.flavor, width=kn.width, height=kn.height, depth=kn.depth, sample_gap=kn.sample_gap,
                      text_length=kn.text_length, max_retain_ratio=kn.max_retain_ratio,
                      min_retain_ratio=kn.min_retain_ratio, estimator="meanpool")
    qs, ks, vs = (x[:, :, :heads].transpose(1, 2) for x in (q, k, v))          # [B,heads,S,D] views
    rr = O.GilbertRearranger(cfg.width, cfg.height, cfg.depth, cfg.text_length)
    t0 = time.perf_counter()
    res = O.asa_forward(qs, ks, vs, cfg, rearranger=rr)
    dt = time.perf_counter() - t0
    S, D = qs.shape[2], qs.shape[3]
    n_pool = -(-S // cfg.sample_gap)
    cnt = res.mask.sum(-1)
    fl = algorithmic_flops(cnt, res.mask[..., -1], S, D, n_pool)
    return dt, fl, threads


def run_reference(args):
    from video_blade_b200.asa import AsaKnobs
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kn, H, D, wname = workload(args)
    S = kn.width * kn.height * kn.depth + kn.text_length
    q, k, v = make_inputs(1, H, S, D, args.inputs, 0, (kn.width, kn.height, kn.depth), kn.text_length)
    heads = 1
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    for _ in range(warm):
        cpu_reference_sample(q, k, v, kn, heads)
    ts, fl, thr = [], 0.0, 0
    for _ in range(steps):
        dt, fl, thr = cpu_reference_sample(q, k, v, kn, heads)
        ts.append(dt)
    t = sum(ts) / len(ts)
    val = fl / t / 1e12
    sample = f"{heads} of {H} heads of the workload (all {S} query rows, dense-masked fp32 + pooled branch + merge)"
    line = {"metric": "ASA sparse-effective attention throughput (whole layer)", "value": val, "unit": "TFLOP/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": t * 1e3, "ms_per_layer_extrapolated": t * 1e3 * H / heads,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 in / fp32 math",
            "data": "synthetic", "config": {"workload": wname, "inputs": args.inputs},
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
def run_blade(args):
    import torch.distributed as dist
    from video_blade_b200 import _lib
    from video_blade_b200.asa import AsaEngine, AsaKnobs
    from video_blade_b200.ulysses import UlyssesGroup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.check(_lib.load().blade_device_check())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kn, H, D, wname = workload(args)
    S = kn.width * kn.height * kn.depth + kn.text_length
    n_pool = -(-S // kn.sample_gap)
    eng = AsaEngine(kn)
    peaks = load_peaks()

    # ---- headline topology: one sequence per rank (CFG / prompt batch split), no collective on the data path
    B_glob, n_groups, P = world, world, 1
    group_id, prank = rank, 0
    ug = None
    Hl = H

    q, k, v = make_inputs(1, H, S, D, args.inputs, group_id, (kn.width, kn.height, kn.depth), kn.text_length)
    hq, hk, hv = (x.pin_memory() for x in (q, k, v))                                           # host [1,S,H,D]
    dq, dk, dv = (x.to(dev) for x in (hq, hk, hv))
    out_host = torch.empty(dq.shape, dtype=dq.dtype).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def layer(xq, xk, xv):
        """[1,S,H,D] device tensors -> [1,S,H,D] attention output (+ cnt)."""
        o, cnt = eng.forward(xq.transpose(1, 2), xk.transpose(1, 2), xv.transpose(1, 2))
        return o.transpose(1, 2), cnt

    # ---- algorithmic FLOPs of my share (from the actual selection)
    o, cnt = layer(dq, dk, dv)
    torch.cuda.synchronize()
    nb = cnt.shape[-1]
    _, dbg = eng.forward(dq.transpose(1, 2), dk.transpose(1, 2), dv.transpose(1, 2), return_debug=True)
    my_flops = algorithmic_flops(dbg["cnt"], dbg["mask"][..., -1], S, D, n_pool)
    fl_t = torch.tensor([my_flops], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(fl_t)
    total_flops = float(fl_t.item())
    retained_mean = float(dbg["cnt"].float().mean())

    # ---- stage events inside the timed region (attention kernel = dominant kernel)
    lib = _lib.load()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(4)]
          for _ in range(args.steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, fn, with_stage_events=False):
        t0s, t1s = [], []
        for i in range(nsteps):
            flush.zero_()                                # L2 flush between timed iterations (untimed)
            if with_stage_events:
                for s in range(4):
                    ev[i][s][0].record(); ev[i][s][1].record()   # create handles
                    lib.blade_profile_events(s, ev[i][s][0].cuda_event, ev[i][s][1].cuda_event)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            t0s.append(e0); t1s.append(e1)
        torch.cuda.synchronize()
        for s in range(4):
            lib.blade_profile_events(s, None, None)
        return [a.elapsed_time(b) for a, b in zip(t0s, t1s)]

    for _ in range(args.warmup):
        layer(dq, dk, dv)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    wall0 = time.perf_counter()
    per_step = timed(args.steps, lambda: layer(dq, dk, dv), with_stage_events=True)
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(per_step)
    stage_ms = [sum(ev[i][s][0].elapsed_time(ev[i][s][1]) for i in range(args.steps)) / args.steps for s in range(4)]

    # ---- e2e: host buffers; every step copies its q,k,v from pinned host memory and reads its output back.
    # The three stages run on three streams with double-buffered device tensors (step i+1's H2D and step i-1's D2H
    # overlap step i's kernels), the way a serving loop would drive the C ABI; all copies are inside the timed region.
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    s_main = torch.cuda.current_stream()
    dbuf = [[torch.empty_like(dq) for _ in range(3)] for _ in range(2)]
    obuf = [torch.empty_like(out_host) for _ in range(2)]
    ohost = [torch.empty(dq.shape, dtype=dq.dtype).pin_memory() for _ in range(2)]

    def e2e_run(nsteps):
        ev_in = [torch.cuda.Event() for _ in range(nsteps)]
        ev_cmp = [torch.cuda.Event() for _ in range(nsteps)]
        ev_out = [torch.cuda.Event() for _ in range(nsteps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_main)
        s_in.wait_event(e0); s_out.wait_event(e0)
        for i in range(nsteps):
            slot = i & 1
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_cmp[i - 2])                     # slot's previous consumer is done
                for dst, src in zip(dbuf[slot], (hq, hk, hv)):
                    dst.copy_(src, non_blocking=True)
                ev_in[i].record(s_in)
            s_main.wait_event(ev_in[i])
            if i >= 2:
                s_main.wait_event(ev_out[i - 2])                       # output slot drained
            oo, _ = layer(*dbuf[slot])
            obuf_i = oo                                                # [1,S,H,D] view of fresh memory
            ev_cmp[i].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[i])
                ohost[slot].copy_(obuf_i, non_blocking=True)
                obuf_i.record_stream(s_out)
                ev_out[i].record(s_out)
        s_main.wait_stream(s_in); s_main.wait_stream(s_out)
        e1.record(s_main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    e2e_run(min(3, args.steps))
    barrier()
    e2e_ms = e2e_run(args.steps)
    barrier()
    out_host = ohost[0]
    clk = clocks.stop() if rank == 0 else None      # sampled over the timed layer loop and the e2e loop

    # ---- Ulysses configuration (config 3 topology): B = 2 as 2 CFG groups x Ulysses N/2
    ulysses = None
    if world > 1 and world % 2 == 0 and H % (world // 2) == 0 and S % (world // 2) == 0:
        Pu = world // 2
        ugrp = UlyssesGroup(world, rank, Pu)
        gid, pr = rank // Pu, rank % Pu
        uq, uk, uv = make_inputs(1, H, S, D, args.inputs, gid, (kn.width, kn.height, kn.depth), kn.text_length)
        sl = slice(pr * (S // Pu), (pr + 1) * (S // Pu))
        uq, uk, uv = (x[:, sl].contiguous().to(dev) for x in (uq, uk, uv))                    # my sequence shard

        def ulayer(ev=None):
            if ev: ev[0].record()
            gq, gk, gv, vrow, _keep = ugrp.scatter_heads_fused(uq, uk, uv)                      # one all_to_all, no unpack
            if ev: ev[1].record()
            o, _ = eng.forward(gq, gk, gv, virtual_rows=vrow)
            if ev: ev[2].record()
            r = ugrp.gather_heads(o.transpose(1, 2))                                           # [1,S/P,H,D]
            if ev: ev[3].record()
            return r
        for _ in range(args.warmup):
            ulayer()
        barrier()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
        for i in range(args.steps):
            flush.zero_()
            ulayer(evs[i])
        barrier()
        seg = [sum(evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(args.steps)) / args.steps for j in range(3)]
        tt = torch.tensor([sum(seg)] + seg, dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tot, sc, fw, ga = (float(x) for x in tt.tolist())
        pair_flops = total_flops / world * 2                                                   # two sequences
        ulysses = {"parallelism": f"cfg2xulysses{Pu}", "global_batch": 2, "ms_per_layer": tot,
                   "scatter_all_to_all_ms": sc, "asa_ms": fw, "gather_all_to_all_ms": ga,
                   "value": pair_flops / (tot * 1e-3) / 1e12, "unit": "TFLOP/s", "scaling": "strong",
                   "all_to_all_bytes_per_rank": int((3 + 1) * uq.numel() * 2 * (Pu - 1) / Pu)}

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = (float(x) for x in t.tolist())
    ms_per_step = dev_ms / args.steps
    value = total_flops / (ms_per_step * 1e-3) / 1e12
    e2e_value = total_flops / (e2e_ms / args.steps * 1e-3) / 1e12

    if rank == 0:
        attn_ms = stage_ms[3]
        attn_flops = my_flops                          # this rank's launch
        achieved = attn_flops / (attn_ms * 1e-3) / 1e12
        line = {
            "metric": "ASA sparse-effective attention throughput (whole layer)", "value": value, "unit": "TFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "ms_per_layer": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wname, "inputs": args.inputs, "global_batch": B_glob,
                       "parallelism": "single" if world == 1 else f"batch{world} (CFG/prompt split, no data-path collective)",
                       "l2": "256 MiB flush between timed iterations", "retained_blocks_per_row_mean": retained_mean,
                       "algorithmic_tflop_per_step": total_flops / 1e12,
                       "stage_ms": {"prep": stage_ms[0], "scores": stage_ms[1], "select": stage_ms[2],
                                    "attention": stage_ms[3]}},
            "e2e": {"value": e2e_value, "unit": "TFLOP/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": 3 * hq.numel() * 2 * world, "d2h_bytes_per_step": out_host.numel() * 2 * world},
            "gpu_launches": args.steps * 5,
            "roofline": {"kernel": f"asa_attn_kernel<{D},bf16>", "bound": "tensor", "achieved": achieved,
                         "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"],
                         "frac_of_sustained": achieved / peaks["tf_sustained"] if peaks["tf_sustained"] else None,
                         "peak_source": peaks["source"], "avg_launch_ms": attn_ms,
                         "traffic": traffic_bytes() if args.workload == "wan" and args.retain is None else None,
                         "traffic_unit": "bytes/launch (ncu dram read+write, profiles/traffic.json)"},
            "roofline_maskgen": {"kernel": "prep_block_kernel (gather + copy + block means)", "bound": "hbm",
                                 "achieved": (2 * 3 * Hl * S * D * 2) / (stage_ms[0] * 1e-3) / 1e9,
                                 "peak": peaks["hbm"], "unit": "GB/s", "avg_ms": stage_ms[0]},
            "clocks": clk, "wall_s_timed_region": wall,
        }
        line["roofline_maskgen"]["frac"] = line["roofline_maskgen"]["achieved"] / peaks["hbm"]
        if ulysses:
            line["ulysses"] = ulysses
        if world == 1 and not args.no_cpu_baseline:
            cpu_heads = min(4, Hl)                 # ~10 s of host work on a 16-core box
            dt, fl, thr = cpu_reference_sample(q, k, v, kn, heads=cpu_heads)
            line["cpu_baseline"] = {"value": fl / dt / 1e12, "unit": "TFLOP/s", "cores": thr, "kind": "port",
                                    "seconds": dt,
                                    "sample": f"{cpu_heads} of {H} heads of the workload (all {S} query rows, dense-masked "
                                              "fp32 + pooled branch + merge), one pass"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="blade", choices=["blade", "reference"])
    ap.add_argument("--inputs", default="gaussian", choices=["gaussian", "structured"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="wan", choices=["wan", "cog"],
                    help="wan = BASELINE config 2 (default, the headline); cog = config 4")
    ap.add_argument("--retain", type=float, default=None, help="config 5 density sweep: min = max retain ratio")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "blade" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_blade(args)


if __name__ == "__main__":
    main()
