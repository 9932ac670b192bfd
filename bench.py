#!/usr/bin/env python
"""bench.py -- the driver contract for the ASA hot path (BASELINE.json: "ASA ms/layer + sparse-eff TFLOP/s
@Wan 32760 tok; 8-step clip s at 1/2/4/8 GPU").

One "step" = one Adaptive-Sparse-Attention layer call (`AdaptiveBlockSparseAttnTrain.forward`, W:383-408) on synthetic
bf16 q,k,v of BASELINE config 2: Wan2.1-T2V-1.3B, 81x480x832 -> [B,12,32760,128], Gilbert rearrangement on, block 128,
retain 5-17 %, energy 0.95, pooled branch gap 30.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl blade|reference] [--inputs gaussian|structured|mixed]

N = 1 : the workload is exactly config 2 (B = 1).
N > 1 : torchrun, one rank per GPU over NCCL.  The workload is FIXED as N grows (STRONG scaling): the classifier-free-
        guidance pair of config 3 (B = 2 sequences) as 2 CFG groups x Ulysses degree N/2 (Wan: 12 heads; N = 2 -> 2 x 1,
        4 -> 2 x 2, 8 -> 2 x 4).  `value` = algorithmic FLOPs of the two sequences' layer / the max-over-ranks device time
        of the sharded layer INCLUDING the head/sequence exchange on both sides.  Data plane (config.data_plane):
        "p2p" = the gather kernel pulls q/k/v rows from the peers and the attention epilogue pushes output rows to the
        peers over NVLink peer memory (no all-to-all launches); "nccl" = two all_to_all_single around the layer.
        Rank outputs are checked against the single-GPU layer on the same sequence (`ulysses_parity`).  The old
        replica figure (one sequence per rank, no exchange) is reported under "replicas".
`value` = algorithmic sparse-attention FLOPs per second, inputs resident in HBM.
`e2e`   = the same metric through the host-facing call: pinned host q,k,v -> H2D -> layer -> D2H, every step.
`clip`  = BASELINE's second metric: seconds for an 8-step CFG clip of the Wan-shaped 30-block DiT at N GPUs.

--impl reference times the reference's own CPU path (dense-masked PyTorch substitute for the absent CUDA library; the
oracle port, since the Python reference tree does not travel to the GPU box) on the host cores, on a bounded sample of
the same workload.
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
METRIC = "ASA sparse-effective attention throughput (whole layer)"
STAGES = ("prep", "scores", "select", "attention", "pool")


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
    if args.estimator != "meanpool":
        kn.estimator = args.estimator
        name += f", estimator {args.estimator}"
    return kn, H, D, name


def base_config(args, wname):
    """Keys shared by the GPU arm and the reference arm (the driver compares the two `config` dicts)."""
    return {"workload": wname, "inputs": args.inputs, "estimator": args.estimator}


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


def traffic_bytes(D):
    """ncu dram bytes per launch of the attention kernel, from this round's capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)[f"asa_attn_kernel<{D},bf16>"]["dram_bytes_per_launch"]
    except Exception:
        return None


def bind_to_gpu_numa_node(local):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off before the pinned host buffers are allocated
    (first touch places them there): the e2e leg is PCIe/host bound and a remote node halves the copy rate."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return {"node": node, "cpus": len(ids)}
    except Exception:
        return None
    return None


def make_inputs(B, H, S, D, kind, seed, grid, text_length=0):
    from video_blade_b200.synth import synth_qkv     # synthetic-input recipe shared with the tests (not timed)
    amp, ramp = {"gaussian": (0.0, False), "structured": (2.0, False), "mixed": (2.5, True)}[kind]
    q, k, v = synth_qkv(B, H, S, D, seed=seed, structured=amp, grid=grid, text_length=text_length, ramp=ramp)
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


def maskgen_bytes(B, H, S, D, nb):
    """read Q + read K once + write the index list (+ counts) + write the fp32 scores."""
    return 2 * B * H * S * D * 2 + B * H * nb * (nb + 1) * 4 + B * H * nb * nb * 4


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_sample(q, k, v, kn, heads=1, threads=None, estimator="meanpool"):
    """The reference's CPU path (oracle port: reference module code + dense-masked substitute) on `heads`
    heads of the workload.  Returns (seconds, algorithmic FLOPs of the sample, threads)."""
    from oracle import asa_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = O.ASAConfig(flavor=kn.flavor, width=kn.width, height=kn.height, depth=kn.depth, sample_gap=kn.sample_gap,
                      text_length=kn.text_length, max_retain_ratio=kn.max_retain_ratio,
                      min_retain_ratio=kn.min_retain_ratio, estimator=estimator)
    qs, ks, vs = (x[:, :, :heads].transpose(1, 2) for x in (q, k, v))          # [B,heads,S,D] views
    rr = O.GilbertRearranger(cfg.width, cfg.height, cfg.depth, cfg.text_length)
    qo = ko = None
    if estimator == "sampled_max":
        g = torch.Generator().manual_seed(0)
        qo = O.draw_sample_offsets(1, heads, 128, 32, g)
        ko = O.draw_sample_offsets(1, heads, 128, 32, g)
    t0 = time.perf_counter()
    res = O.asa_forward(qs, ks, vs, cfg, qo, ko, rearranger=rr)
    dt = time.perf_counter() - t0
    S, D = qs.shape[2], qs.shape[3]
    n_pool = -(-S // cfg.sample_gap)
    cnt = res.mask.sum(-1)
    fl = algorithmic_flops(cnt, res.mask[..., -1], S, D, n_pool)
    return dt, fl, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kn, H, D, wname = workload(args)
    S = kn.width * kn.height * kn.depth + kn.text_length
    q, k, v = make_inputs(1, H, S, D, args.inputs, 0, (kn.width, kn.height, kn.depth), kn.text_length)
    heads = 2
    steps = max(1, min(args.steps, 3))          # each step is ~10-20 s of host work: bounded so the run ends in minutes
    warm = 1 if args.warmup > 0 else 0
    for _ in range(warm):
        cpu_reference_sample(q, k, v, kn, heads, estimator=args.estimator)
    ts, fl, thr = [], 0.0, 0
    for _ in range(steps):
        dt, fl, thr = cpu_reference_sample(q, k, v, kn, heads, estimator=args.estimator)
        ts.append(dt)
    t = sum(ts) / len(ts)
    val = fl / t / 1e12
    sample = (f"{heads} of {H} heads of the workload (all {S} query rows, dense-masked fp32 + pooled branch + merge); "
              f"steps capped at 3; whole-layer time is EXTRAPOLATED x{H // heads} (heads are independent)")
    cfg = base_config(args, wname)
    line = {"metric": METRIC, "value": val, "unit": "TFLOP/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": t * 1e3, "ms_per_layer_extrapolated": t * 1e3 * H / heads,
            "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "bf16 in / fp32 math", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
def run_blade(args):
    import torch.distributed as dist
    from video_blade_b200 import _lib
    from video_blade_b200.asa import AsaEngine
    from video_blade_b200.ulysses import UlyssesGroup, UlyssesPeerPlane

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    _lib.check(_lib.load().blade_device_check())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kn, H, D, wname = workload(args)
    S = kn.width * kn.height * kn.depth + kn.text_length
    nb = -(-S // kn.block_size)
    n_pool = -(-S // kn.sample_gap)
    eng = AsaEngine(kn)
    peaks = load_peaks()
    lib = _lib.load()
    grid3 = (kn.width, kn.height, kn.depth)

    # topology: N = 1 -> one sequence; N > 1 -> the CFG pair as 2 groups x Ulysses N/2 (strong scaling)
    if world == 1:
        n_seq, P, gid, pr = 1, 1, 0, 0
    else:
        assert world % args.groups == 0, "N must be a multiple of --groups"
        n_seq, P = args.groups, world // args.groups
        gid, pr = rank // P, rank % P
        assert H % P == 0 and S % P == 0, f"Ulysses degree {P} does not divide heads {H} / tokens {S}"
    Sl, Hl = S // P, H // P

    q, k, v = make_inputs(1, H, S, D, args.inputs, gid, grid3, kn.text_length)                 # my group's sequence
    hq, hk, hv = (x.pin_memory() for x in (q, k, v))                                           # host [1,S,H,D]
    dq, dk, dv = (x.to(dev) for x in (hq, hk, hv))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    sample_off = None
    if kn.estimator == "sampled_max":
        g = torch.Generator(device=dev).manual_seed(7)
        sample_off = (eng.draw_offsets(1, H, dev, g), eng.draw_offsets(1, H, dev, g))

    def layer(xq, xk, xv, **kw):
        """[1,S,H,D] device tensors -> [1,S,H,D] attention output (+ cnt)."""
        o, cnt = eng.forward(xq.transpose(1, 2), xk.transpose(1, 2), xv.transpose(1, 2), sample_offsets=sample_off, **kw)
        return o.transpose(1, 2), cnt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- algorithmic FLOPs of one sequence (from the actual selection), and the single-GPU output for the parity check
    ref_out, dbg = layer(dq, dk, dv, return_debug=True)
    torch.cuda.synchronize()
    seq_flops = algorithmic_flops(dbg["cnt"], dbg["mask"][..., -1], S, D, n_pool)
    cnt_all = dbg["cnt"].float()
    retained = {"mean": float(cnt_all.mean()), "min": int(cnt_all.min()), "max": int(cnt_all.max())}
    fl_t = torch.tensor([seq_flops if pr == 0 else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(fl_t)
    total_flops = float(fl_t.item())                 # N = 1: one sequence; N > 1: the two sequences of the pair

    # ---- section A: the single-GPU layer on this rank's sequence, with the per-stage events
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in STAGES]
          for _ in range(args.steps)]

    def timed(nsteps, fn, with_stage_events=False):
        t0s, t1s = [], []
        for i in range(nsteps):
            flush.zero_()                                # L2 flush between timed iterations (untimed)
            if with_stage_events:
                for s in range(len(STAGES)):
                    ev[i][s][0].record(); ev[i][s][1].record()   # create handles
                    lib.blade_profile_events(s, ev[i][s][0].cuda_event, ev[i][s][1].cuda_event)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            t0s.append(e0); t1s.append(e1)
        torch.cuda.synchronize()
        for s in range(len(STAGES)):
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
    single_ms = sum(per_step) / args.steps
    stage_ms = {n: sum(ev[i][s][0].elapsed_time(ev[i][s][1]) for i in range(args.steps)) / args.steps
                for s, n in enumerate(STAGES)}

    # model-level hoist of the Gilbert permutation: tokens already in curve order, the layer neither
    # gathers nor un-permutes -- the mask-generation front end then reads Q and K exactly once
    hoisted = None
    if world == 1 and kn.estimator == "meanpool":
        from dataclasses import replace
        eng_h = AsaEngine(replace(kn, use_rearrange=False))
        src = eng.src_row(dev, S).long()
        pq, pk, pv = (x[:, src].contiguous() for x in (dq, dk, dv))

        def layer_h():
            return eng_h.forward(pq.transpose(1, 2), pk.transpose(1, 2), pv.transpose(1, 2))
        o_h, _ = layer_h()
        back = torch.empty_like(o_h.transpose(1, 2))
        back[:, src] = o_h.transpose(1, 2)
        hoist_equal = bool(torch.equal(back, ref_out))
        for _ in range(args.warmup):
            layer_h()
        torch.cuda.synchronize()
        hs = timed(args.steps, layer_h, with_stage_events=True)
        h_stage = {n: sum(ev[i][s][0].elapsed_time(ev[i][s][1]) for i in range(args.steps)) / args.steps
                   for s, n in enumerate(STAGES)}
        chain = h_stage["prep"] + h_stage["scores"] + h_stage["select"]
        mb = maskgen_bytes(1, H, S, D, nb)
        hoisted = {"ms_per_layer": sum(hs) / args.steps, "stage_ms": h_stage, "bit_equal_to_per_layer_gather": hoist_equal,
                   "value": seq_flops / (sum(hs) / args.steps * 1e-3) / 1e12, "unit": "TFLOP/s",
                   "roofline_maskgen": {"bound": "hbm", "bytes": mb, "chain_ms": chain,
                                        "achieved": mb / (chain * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                        "frac": mb / (chain * 1e-3) / 1e9 / peaks["hbm"]}}
        del pq, pk, pv, eng_h

    # ---- section C (N > 1): the CFG x Ulysses pair, device-resident inputs
    ulysses, pair_ms, plane_name, parity = None, None, "none", None
    plane = None
    if world > 1:
        ugrp = UlyssesGroup(world, rank, P)
        sl = slice(pr * Sl, (pr + 1) * Sl)
        uq, uk, uv = (x[:, sl].contiguous() for x in (dq, dk, dv))                             # my token shard [1,Sl,H,D]
        modes = {}
        if P == 1:
            def run_pair():
                return layer(dq, dk, dv)[0][0]                                                 # [S,H,D]
            modes["none"] = run_pair
        else:
            def run_nccl():
                gq, gk, gv, vrow, _keep = ugrp.scatter_heads_fused(uq, uk, uv)                  # one all_to_all, no unpack
                o, _ = eng.forward(gq, gk, gv, virtual_rows=vrow, sample_offsets=None if sample_off is None else
                                   tuple(x[:, pr * Hl:(pr + 1) * Hl].contiguous() for x in sample_off))
                return ugrp.gather_heads(o.transpose(1, 2))[0]                                 # [Sl,H,D]
            modes["nccl"] = run_nccl
            if not args.no_p2p:
                try:
                    plane = UlyssesPeerPlane(ugrp, Sl, H, D, dtype=dq.dtype, device=dev)
                    for j, x in enumerate((uq, uk, uv)):
                        plane.qkv[j].copy_(x[0])                                               # "projection outputs" in place

                    def run_p2p():
                        so = None if sample_off is None else tuple(x[:, pr * Hl:(pr + 1) * Hl].contiguous() for x in sample_off)
                        o, _ = plane.attention(eng, sample_offsets=so)
                        return o
                    run_p2p()
                    torch.cuda.synchronize()
                    modes["p2p"] = run_p2p
                except Exception as e:                                                          # no peer mapping on this box
                    plane = None
                    modes_err = f"{type(e).__name__}: {e}"
                    if rank == 0:
                        print(f"[bench] peer-memory plane unavailable, NCCL only: {modes_err}", file=sys.stderr)
        # every rank must agree on the mode list (a failed rendezvous on one rank disables p2p everywhere)
        have_p2p = torch.tensor([1 if "p2p" in modes else 0], device=dev)
        dist.all_reduce(have_p2p, op=dist.ReduceOp.MIN)
        if not int(have_p2p.item()):
            modes.pop("p2p", None)
        want = ref_out[0, sl] if P > 1 else ref_out[0]
        results = {}
        for name, fn in modes.items():
            got = fn()
            torch.cuda.synchronize()
            d = got.float() - want.float()
            par = torch.tensor([float(torch.equal(got, want)), float(d.norm()) ** 2, float(want.float().norm()) ** 2,
                                float(d.abs().max())], dtype=torch.float64, device=dev)
            mx = par[3:].clone()
            dist.all_reduce(par[:1], op=dist.ReduceOp.MIN)
            dist.all_reduce(par[1:3])
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            for _ in range(args.warmup):
                fn()
            barrier()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for i in range(args.steps):
                flush.zero_()
                evs[i][0].record()
                fn()
                evs[i][1].record()
            barrier()
            tt = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / args.steps], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            results[name] = {"ms_per_layer_pair": float(tt.item()),
                             "parity": {"bit_exact": bool(par[0].item() == 1.0),
                                        "rel_l2": float((par[1] / par[2]).sqrt().item()), "max_abs": float(mx.item())}}
        # stage breakdown of the peer-plane layer on this rank (events inside the C call + around the two barriers)
        if "p2p" in modes:
            sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in STAGES]
            bev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            acc = {n: 0.0 for n in STAGES}
            acc.update({"barrier_in": 0.0, "barrier_out": 0.0, "total": 0.0})
            for i in range(args.steps):
                flush.zero_()
                for s_i in range(len(STAGES)):
                    sev[s_i][0].record(); sev[s_i][1].record()
                    lib.blade_profile_events(s_i, sev[s_i][0].cuda_event, sev[s_i][1].cuda_event)
                bev[0].record()
                plane.barrier()
                bev[1].record()
                q_, k_, v_ = plane.views()
                pe_, out_ = plane.launch_args()
                eng.forward(q_, k_, v_, peers=pe_, out=out_)
                bev[2].record()
                plane.barrier()
                bev[3].record()
                torch.cuda.synchronize()
                for s_i, n in enumerate(STAGES):
                    acc[n] += sev[s_i][0].elapsed_time(sev[s_i][1])
                acc["barrier_in"] += bev[0].elapsed_time(bev[1])
                acc["barrier_out"] += bev[2].elapsed_time(bev[3])
                acc["total"] += bev[0].elapsed_time(bev[3])
            for s_i in range(len(STAGES)):
                lib.blade_profile_events(s_i, None, None)
            results["p2p"]["rank0_stage_ms"] = {k_: v_ / args.steps for k_, v_ in acc.items()}
        plane_name = min(results, key=lambda n: results[n]["ms_per_layer_pair"])
        pair_ms = results[plane_name]["ms_per_layer_pair"]
        parity = results[plane_name]["parity"]
        shard_bytes = Sl * H * D * 2
        ulysses = {"parallelism": f"cfg{n_seq}xulysses{P}", "global_batch": n_seq, "data_planes": results, "chosen": plane_name,
                   "bytes_exchanged_per_rank": int((3 + 1) * shard_bytes * (P - 1) / P),
                   "single_gpu_pair_ms": n_seq * single_ms}

    # ---- section B: e2e through host buffers; every step copies its q,k,v from pinned host memory and reads its
    # output back.  The three stages run on three streams with double-buffered device tensors (step i+1's H2D and
    # step i-1's D2H overlap step i's kernels), the way a serving loop would drive the C ABI.
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    s_main = torch.cuda.current_stream()
    use_plane = world > 1 and P > 1
    if use_plane:                      # my token shard of the pair
        sl = slice(pr * Sl, (pr + 1) * Sl)
        host_in = [x[:, sl].contiguous().pin_memory() for x in (q, k, v)]
    else:
        host_in = [hq, hk, hv]
    in_shape = host_in[0].shape
    dbuf = [[torch.empty(in_shape, dtype=dq.dtype, device=dev) for _ in range(3)] for _ in range(2)]
    ohost = [torch.empty(in_shape, dtype=dq.dtype).pin_memory() for _ in range(2)]

    def e2e_layer(bufs):
        if not use_plane:
            return layer(*bufs)[0]
        if plane_name == "p2p":
            for j in range(3):
                plane.qkv[j].copy_(bufs[j][0])           # the projection GEMMs would write here directly
            return plane.attention(eng)[0].unsqueeze(0)
        return _nccl_from(bufs)

    def _nccl_from(bufs):
        gq, gk, gv, vrow, _keep = ugrp.scatter_heads_fused(*bufs)
        o, _ = eng.forward(gq, gk, gv, virtual_rows=vrow)
        return ugrp.gather_heads(o.transpose(1, 2))

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
                for dst, src in zip(dbuf[slot], host_in):
                    dst.copy_(src, non_blocking=True)
                ev_in[i].record(s_in)
            s_main.wait_event(ev_in[i])
            if i >= 2:
                s_main.wait_event(ev_out[i - 2])                       # output slot drained
            oo = e2e_layer(dbuf[slot])
            if use_plane and plane_name == "p2p":
                oo = oo.clone()                                        # the symmetric buffer is rewritten next step
            ev_cmp[i].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[i])
                ohost[slot].copy_(oo.reshape(in_shape), non_blocking=True)
                oo.record_stream(s_out)
                ev_out[i].record(s_out)
        s_main.wait_stream(s_in); s_main.wait_stream(s_out)
        e1.record(s_main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    e2e_run(min(3, args.steps))
    barrier()
    e2e_ms = e2e_run(args.steps)
    barrier()
    clk = clocks.stop() if rank == 0 else None      # sampled over the timed layer loop, the pair and the e2e loop

    t = torch.tensor([single_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    single_ms_max, e2e_ms = (float(x) for x in t.tolist())

    # ---- BASELINE's second metric: the 8-step clip at this N
    clip = None
    if not args.no_clip and args.workload == "wan" and args.retain is None:
        del dbuf, flush
        torch.cuda.empty_cache()
        try:
            import bench_clip
            clip = bench_clip.run_clip(world, rank, dev, steps=8, layers=30, reps=1, model="wan")
        except Exception as e:
            clip = {"error": f"{type(e).__name__}: {e}"}

    if world == 1:
        ms_per_step = single_ms_max
    else:
        ms_per_step = pair_ms
    value = total_flops / (ms_per_step * 1e-3) / 1e12
    e2e_value = total_flops / (e2e_ms / args.steps * 1e-3) / 1e12

    if rank == 0:
        attn_ms = stage_ms["attention"]
        achieved = seq_flops / (attn_ms * 1e-3) / 1e12            # this rank's single-GPU launch: one full sequence
        chain_ms = stage_ms["prep"] + stage_ms["scores"] + stage_ms["select"]
        mb = maskgen_bytes(1, H, S, D, nb)
        n_kernels = 5 + (2 if kn.estimator == "sampled_max" else 0)
        cfg = base_config(args, wname)
        cfg.update({"global_batch": n_seq,
                    "parallelism": "single" if world == 1 else f"cfg{n_seq}xulysses{P}",
                    "data_plane": plane_name,
                    "l2": "256 MiB flush between timed iterations", "retained_blocks_per_row": retained,
                    "algorithmic_tflop_per_step": total_flops / 1e12,
                    "stage_ms": stage_ms, "host_numa_binding": numa})
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "ms_per_layer": single_ms_max if world == 1 else pair_ms / 2,
            "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": "TFLOP/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": 3 * S * H * D * 2 * n_seq, "d2h_bytes_per_step": S * H * D * 2 * n_seq},
            "gpu_launches": args.steps * n_kernels,
            "roofline": {"kernel": f"asa_attn_kernel<{D},bf16>", "bound": "tensor", "achieved": achieved,
                         "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"],
                         "frac_of_sustained": achieved / peaks["tf_sustained"] if peaks["tf_sustained"] else None,
                         "peak_source": peaks["source"], "avg_launch_ms": attn_ms,
                         "measured_on": "the single-GPU launch of one full sequence (section A)",
                         "traffic": traffic_bytes(D) if args.retain is None and args.inputs == "gaussian" else None,
                         "traffic_unit": "bytes/launch (ncu dram read+write, profiles/traffic.json)"},
            "roofline_maskgen": {"kernels": "prep_block (gather + block means) + score + select, whole chain",
                                 "bound": "hbm", "bytes": mb, "bytes_definition": "SURVEY 8(d): Q + K read once + index "
                                 "list + fp32 scores", "chain_ms": chain_ms,
                                 "achieved": mb / (chain_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                 "frac": mb / (chain_ms * 1e-3) / 1e9 / peaks["hbm"],
                                 "note": "the drop-in layer also gathers q,k,v into curve order here (reference semantics: "
                                         "permute per layer, W:142-159); `hoisted` is the model-level-permutation mode"},
            "clocks": clk, "wall_s_timed_region": wall,
        }
        if hoisted:
            line["hoisted"] = hoisted
        if ulysses:
            line["ulysses"] = ulysses
            line["ulysses_parity"] = parity
            line["replicas"] = {"ms_per_layer": single_ms_max, "value": seq_flops * world / (single_ms_max * 1e-3) / 1e12,
                                "unit": "TFLOP/s", "note": "one sequence per rank, no exchange (weak scaling, round-1 headline)"}
        if clip is not None:
            line["clip"] = clip
        if world == 1 and not args.no_cpu_baseline:
            cpu_heads = min(4, H)                 # ~10 s of host work on a 16-core box
            dt, fl, thr = cpu_reference_sample(q, k, v, kn, heads=cpu_heads, estimator=args.estimator)
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
    ap.add_argument("--inputs", default="gaussian", choices=["gaussian", "structured", "mixed"])
    ap.add_argument("--estimator", default="meanpool", choices=["meanpool", "sampled_max"],
                    help="meanpool = north-star kernel (a) (default); sampled_max = the reference's estimator (W:62-87)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clip", action="store_true", help="skip the 8-step clip (BASELINE metric part 2)")
    ap.add_argument("--no-p2p", action="store_true", help="N > 1: NCCL all-to-all data plane only")
    ap.add_argument("--groups", type=int, default=2,
                    help="N > 1: number of independent sequences (CFG branches); Ulysses degree = N / groups.  The "
                         "contract run uses 2 (the CFG pair); 1 is a diagnostic (one sequence over all N GPUs)")
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
