"""Quick timing of the train step and the embedding pass (device-resident inputs)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hippie_b200.engine import Engine

B = int(os.environ.get("B", "512"))
steps = int(os.environ.get("STEPS", "20"))
eng = Engine(10, 50, 100, 5, 5, 5, True, B, False, int(os.environ.get("CONV_PATH", "0"))).allocate("cuda:0")
g = torch.Generator().manual_seed(0)
torch.manual_seed(42)
from hippie_b200.model import MultiModalCVAE
m = MultiModalCVAE(10, 50, 100, 5, 5, 5, max_batch=B)
eng.flat_params.copy_(m._flat["params"])
dev = torch.device("cuda:0")
x1 = (0.365 * torch.randn(B, 1, 50, generator=g) + 0.019).clamp(-1, 1.3).to(dev)
x2 = torch.log1p(0.0157 * torch.randn(B, 1, 100, generator=g).abs()).to(dev)
src = torch.randint(1, 5, (B,), generator=g).to(dev)
eps = torch.randn(B, 10, generator=g).to(dev)
scal = torch.zeros(8, device=dev)
def step(i):
    eng.train_fwd_bwd(x1, x2, src, None, eps, 0.5, 1.0, 1.0, scalars=scal)
    eng.clip_adamw(1e-3, 0.01, i + 1, max_norm=1.0, scalars=scal)
for i in range(5):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for i in range(steps):
    step(5 + i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"train B={B}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} samples/s  (host wall {1e3 * (time.time() - t0) / steps:.3f} ms/step)  loss {scal[0].item():.5f} gnorm {scal[4].item():.4f}")
print("launches per step", eng.last_launch_count() + 3, " TFLOP/s (algorithmic 689.76 MFLOP/sample): %.2f" % (B * 689.76456e6 / (ms * 1e-3) / 1e12))
for i in range(3):
    eng.embed(x1, x2, src, None)
torch.cuda.synchronize()
e0.record()
for i in range(steps):
    eng.embed(x1, x2, src, None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"embed B={B}: {ms:.3f} ms  {B / ms * 1e3:.0f} samples/s")
