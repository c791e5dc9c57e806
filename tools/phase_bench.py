"""Times the phases of the step separately (forward only / forward + backward / optimizer) with CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
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


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


step = [0]
def opt():
    step[0] += 1
    eng.clip_adamw(1e-3, 0.01, step[0], max_norm=1.0, scalars=scal)

print(f"B={B}")
print(f"train forward only      {timeit(lambda: eng.train_forward(x1, x2, src, None, eps, 0.5, 1.0, 1.0, scalars=scal)):.3f} ms")
print(f"eval forward            {timeit(lambda: eng.eval_forward(x1, x2, src, None, eps, 0.5, 1.0, 1.0, scalars=scal)):.3f} ms")
print(f"embed                   {timeit(lambda: eng.embed(x1, x2, src, None)):.3f} ms")
print(f"train fwd+bwd           {timeit(lambda: eng.train_fwd_bwd(x1, x2, src, None, eps, 0.5, 1.0, 1.0, scalars=scal)):.3f} ms")
print(f"clip+adamw              {timeit(opt):.3f} ms")
