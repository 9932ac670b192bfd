"""Timeline of the attention pipeline from clock64 stamps (needs the -DBLADE_TRACE build, BLADE_ASA_LIB=...)."""
import ctypes as C, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_blade_b200 import _lib
from video_blade_b200.asa import AsaEngine, AsaKnobs
model = sys.argv[1] if len(sys.argv) > 1 else "wan"
kn = AsaKnobs.wan() if model == "wan" else AsaKnobs.cog()
eng = AsaEngine(kn)
S, H, D = (32760, 12, 128) if model == "wan" else (17776, 48, 64)
torch.manual_seed(0)
q, k, v = (torch.randn(1, S, H, D, device="cuda", dtype=torch.bfloat16).transpose(1, 2) for _ in range(3))
for _ in range(2):
    out, cnt = eng.forward(q, k, v)
torch.cuda.synchronize()
buf = np.zeros((4, 8, 256), np.int64)
assert _lib.load().blade_debug_trace(C.c_void_p(buf.ctypes.data)) == 0
t0 = buf[buf > 0].min()
b = np.where(buf > 0, buf - t0, -1)
names = {0: "s_full", 1: "ld_done", 2: "max_done", 3: "exp_done", 4: "arrived"}
for t in (0, 1):
    print(f"== softmax stream {t}: tile: s_full  +ld  +max  +exp  +arrive | gap to next s_full")
    for n in range(8, 40):
        r = b[t, :5, n]
        nxt = b[t, 0, n + 1]
        print(f"  {n:3d}: {r[0]:8d}  {r[1]-r[0]:5d} {r[2]-r[1]:5d} {r[3]-r[2]:5d} {r[4]-r[3]:5d} | total {r[4]-r[0]:5d}  wait_next {nxt-r[4]:5d}  period {nxt-r[0]:5d}")
for t in (0, 1):
    print(f"== mma stream {t}: tile: p_full_seen  v_ready(+)  | qk_issue_of_next(+ from p_full)")
    for n in range(8, 40):
        r = b[2 + t, :3, n]
        print(f"  {n:3d}: {r[0]:8d}  {r[1]-r[0]:5d}  next-QK stamp {b[2+t,2,n+1]-r[0]:6d}   softmax arrive->mma seen {r[0]-b[t,4,n]:5d}   qk_issue->s_full seen {b[t,0,n]-b[2+t,2,n]:5d}")
per = np.diff(b[0, 0, 8:200]); print("stream0 period mean", per.mean(), "median", np.median(per))
for ev in range(4):
    d = (b[0, ev + 1, 8:200] - b[0, ev, 8:200]); print(names[ev + 1], "mean", d.mean())

# item-level view: steps per item = pooled tiles + retained blocks (uniform for Gaussian inputs)
steps = int(-(-(-(-S // kn.sample_gap)) // 128) + int(cnt.flatten()[0]))
print("steps per item", steps)
for t in (0, 1):
    sf = b[t, 0]; ar = b[t, 4]
    n_items = min(6, 250 // steps)
    for it in range(n_items):
        a, z = it * steps, (it + 1) * steps - 1
        inner = np.diff(sf[a:z + 1])
        print(f"stream {t} item {it}: first s_full {sf[a]:8d}  last arrive {ar[z]:8d}  item span {ar[z]-sf[a]:7d}  "
              f"median inner period {int(np.median(inner)):5d}  max inner {inner.max():5d} at step {int(inner.argmax())}  "
              f"gap to next item's first s_full {sf[z+1]-ar[z]:6d}  | after last arrive: o_full seen +{b[t,5,z]-ar[z]:5d}  "
              f"epilogue done +{b[t,6,z]-ar[z]:5d}  next item decoded +{b[t,7,z+1]-ar[z]:5d}  "
              f"mma: next item's first QK issued +{b[2+t,2,z+1]-ar[z]:5d}  epilogue chunks done at "
              + " ".join(f"+{b[2+t,3+c,z]-ar[z]:5d}" for c in range(4)))
