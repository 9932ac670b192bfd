"""Layer time and per-stage split for an arbitrary head count on one GPU (e.g. --heads 3 = the share of one rank of
ulysses 4, without the peer pull): the quick A/B harness for everything that does not shrink with the head count.
    python tools/layer_stages.py --heads 3 [--workload wan|cog] [--inputs gaussian|mixed] [--steps 30]"""
import argparse, importlib.util, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
from video_blade_b200 import _lib
from video_blade_b200.asa import AsaEngine

ap = argparse.ArgumentParser()
ap.add_argument("--heads", type=int, default=3)
ap.add_argument("--workload", default="wan")
ap.add_argument("--inputs", default="gaussian")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--tag", default="")
a = ap.parse_args()
a.retain, a.estimator = None, "meanpool"
kn, _, D, _ = bench.workload(a)
H = a.heads
S = kn.width * kn.height * kn.depth + kn.text_length
eng, lib = AsaEngine(kn), _lib.load()
q, k, v = (x.cuda() for x in bench.make_inputs(1, H, S, D, a.inputs, 0, (kn.width, kn.height, kn.depth), kn.text_length))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
run = lambda: eng.forward(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
for _ in range(5):
    run()
n = a.steps
ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in bench.STAGES] for _ in range(n)]
tt = []
for i in range(n):
    flush.zero_()
    for s in range(len(bench.STAGES)):
        ev[i][s][0].record(); ev[i][s][1].record()
        lib.blade_profile_events(s, ev[i][s][0].cuda_event, ev[i][s][1].cuda_event)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    tt.append((e0, e1))
torch.cuda.synchronize()
for s in range(len(bench.STAGES)):
    lib.blade_profile_events(s, None, None)
ms = sorted(x.elapsed_time(y) for x, y in tt)
st = {nm: round(sum(ev[i][s][0].elapsed_time(ev[i][s][1]) for i in range(n)) / n, 4) for s, nm in enumerate(bench.STAGES)}
print(f"{a.tag} {a.workload} H={H} {a.inputs}: layer median {ms[n // 2]:.4f} ms (min {ms[0]:.4f})  stages {st}", flush=True)
