"""Kernel timeline of the train step (torch.profiler / CUPTI): per-stream busy time, gaps and the critical chain.
  python tools/trace_step.py [out.json]"""
import collections, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from hippie_b200.engine import Engine
from hippie_b200.model import MultiModalCVAE

B = int(os.environ.get("B", "512"))
torch.manual_seed(42)
m = MultiModalCVAE(10, 50, 100, 5, 5, 5, max_batch=B)
eng = Engine(10, 50, 100, 5, 5, 5, True, B).allocate("cuda:0")
eng.flat_params.copy_(m._flat["params"])
g = torch.Generator().manual_seed(0)
dev = eng.device
x1 = (0.365 * torch.randn(B, 1, 50, generator=g) + 0.019).clamp(-1, 1.3).to(dev)
x2 = torch.log1p(0.0157 * torch.randn(B, 1, 100, generator=g).abs()).to(dev)
src = torch.randint(1, 5, (B,), generator=g).to(dev)
eps = torch.randn(B, 10, generator=g).to(dev)
scal = torch.zeros(8, device=dev)
def step(i):
    eng.train_fwd_bwd(x1, x2, src, None, eps, 0.5, 1.0, 1.0, scalars=scal)
    eng.clip_adamw(1e-3, 0.01, i + 1, max_norm=1.0, scalars=scal)
for i in range(6):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        step(6 + i)
    torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "trace_step.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
print(len(ev), "kernel events")
# take the middle step: split at adamw kernels
idx = [i for i, e in enumerate(ev) if "adamw_kernel" in e["name"]]
lo, hi = idx[0] + 1, idx[1] + 1
st = ev[lo:hi]
t0, t1 = st[0]["ts"], max(e["ts"] + e["dur"] for e in st)
print(f"step span {t1 - t0:.1f} us, {len(st)} kernels")
by = collections.defaultdict(list)
for e in st:
    by[e["args"].get("stream")].append(e)
for s, L in sorted(by.items(), key=lambda kv: -len(kv[1])):
    busy = sum(e["dur"] for e in L)
    gaps = [L[i + 1]["ts"] - (L[i]["ts"] + L[i]["dur"]) for i in range(len(L) - 1)]
    pos = [x for x in gaps if x > 0]
    print(f"stream {s}: {len(L)} kernels, busy {busy:.0f} us, span {L[0]['ts'] - t0:.0f}..{L[-1]['ts'] + L[-1]['dur'] - t0:.0f}, "
          f"gaps>0: {len(pos)} sum {sum(pos):.0f} us median {sorted(pos)[len(pos) // 2] if pos else 0:.1f}")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in L:
        mm = re.search(r"(\w+_kernel)(<[^>]*>)?", e["name"])
        k = (mm.group(1) + (mm.group(2) or "")) if mm else e["name"][:40]
        agg[k][0] += 1
        agg[k][1] += e["dur"]
    for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:6]:
        print(f"      {k:42s} {n:4d} {d:8.1f} {d / n:6.1f}")
# concurrency histogram: time with k kernels running
pts = []
for e in st:
    pts.append((e["ts"], 1)), pts.append((e["ts"] + e["dur"], -1))
pts.sort()
cur, last, hist = 0, t0, collections.defaultdict(float)
for t, d in pts:
    hist[cur] += t - last
    cur += d
    last = t
print("time with k kernels in flight:", {k: round(v) for k, v in sorted(hist.items())})
