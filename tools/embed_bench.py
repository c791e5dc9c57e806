import os, sys, torch
sys.path.insert(0, "/root/repo")
from hippie_b200.engine import Engine
B = int(os.environ.get("B", "4096"))
eng = Engine(10, 50, 100, 5, 5, 5, True, B, True).allocate("cuda:0")
torch.manual_seed(0)
eng.flat_params.normal_(0, 0.05)
x1 = torch.randn(B, 1, 50, device="cuda"); x2 = torch.rand(B, 1, 100, device="cuda"); src = torch.randint(1, 5, (B,), device="cuda")
for _ in range(5): eng.embed(x1, x2, src)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): eng.embed(x1, x2, src)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
print(f"embed B={B}: {ms:.3f} ms  {B / ms * 1e3:.0f} samples/s")
