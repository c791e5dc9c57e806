"""One small train step (both label modes) + one embedding pass + the module-level calls, for compute-sanitizer:
    HIPPIE_B200_GRAPHS=0 compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_step.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hippie_b200.model import MultiModalCVAE, MultiModalCVAETrainModule

B = int(os.environ.get("B", "32"))
dev = torch.device("cuda:0")
torch.manual_seed(42)
m = MultiModalCVAE(10, 50, 100, 5, 5, 4, max_batch=B)
tm = MultiModalCVAETrainModule(m, learning_rate=1e-3, weight_decay=0.01, beta=0.5).to(dev)
g = torch.Generator().manual_seed(0)
x1 = (0.365 * torch.randn(B, 1, 50, generator=g) + 0.019).clamp(-1, 1.3).to(dev)
x2 = torch.log1p(0.0157 * torch.randn(B, 1, 100, generator=g).abs()).to(dev)
src = torch.randint(1, 5, (B,), generator=g).to(dev)
cls = torch.randint(0, 4, (B,), generator=g).to(dev)
loss = tm.training_step((x1, x2, src), 0)
tm.optimizer.step(max_norm=1.0)
loss2 = tm.training_step((x1[:B - 3], x2[:B - 3], torch.stack([cls, src], 1)[:B - 3]), 1)  # ragged batch, labelled
tm.optimizer.step(max_norm=1.0)
m.eval()
e = m.embed(x1, x2, src, zscore_ddof=1)
out = m(x1, x2, src)
torch.cuda.synchronize()
print("loss", float(loss), float(loss2), "embed", float(e["enc"].abs().mean()), "flags", m.engine.device_flags())
