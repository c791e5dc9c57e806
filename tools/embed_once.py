"""One embedding pass inside cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import os, sys, torch
sys.path.insert(0, os.environ.get("HIPPIE_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hippie_b200.engine import Engine
B = int(os.environ.get("B", "4096"))
eng = Engine(10, 50, 100, 5, 5, 5, True, B, True).allocate("cuda:0")
torch.manual_seed(0)
eng.flat_params.normal_(0, 0.05)
x1 = torch.randn(B, 1, 50, device="cuda"); x2 = torch.rand(B, 1, 100, device="cuda"); src = torch.randint(1, 5, (B,), device="cuda")
for _ in range(3): eng.embed(x1, x2, src)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.embed(x1, x2, src)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
