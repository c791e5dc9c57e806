"""In-situ per-kernel-class timing of one train step (CUDA events around every launch, single stream, warm L2)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hippie_b200.engine import Engine
from hippie_b200.model import MultiModalCVAE

KINDS = ["conv_fwd", "conv_dgrad", "conv_wgrad", "bn_apply", "bn_bwd(2 launches)", "head_fwd", "head_bwd", "stem_fwd",
         "pool_linear_fwd", "pool_linear_bwd(2)", "stem_wgrad(2)", "dec_linear_fwd", "dec_tail(2)", "dec_linear_bwd(2)",
         "pairsum"]
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
for _ in range(3):
    eng.train_fwd_bwd(x1, x2, src, None, eps, 0.5, 1.0, 1.0, scalars=scal)
torch.cuda.synchronize()
L, h = eng._L, eng._h
L.hippie_profile(h, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.train_fwd_bwd(x1, x2, src, None, eps, 0.5, 1.0, 1.0, scalars=scal)
e1.record()
torch.cuda.synchronize()
print(f"profiled step (single stream, events): {e0.elapsed_time(e1):.3f} ms")
tot = 0.0
for k, name in enumerate(KINDS):
    ms, fl, n = C.c_double(), C.c_double(), C.c_int()
    L.hippie_profile_read(h, k, C.byref(ms), C.byref(fl), C.byref(n))
    if n.value:
        tot += ms.value
        print(f"{name:22s} n={n.value:4d} total {ms.value * 1e3:8.1f} us  avg {ms.value * 1e3 / n.value:7.1f} us")
print(f"sum {tot:.3f} ms")
L.hippie_profile(h, 0)
