"""One bs512 train step inside cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hippie_b200.model import MultiModalCVAE, MultiModalCVAETrainModule

B = int(os.environ.get("B", "512"))
dev = torch.device("cuda:0")
torch.manual_seed(42)
tm = MultiModalCVAETrainModule(MultiModalCVAE(10, 50, 100, 5, 5, 5, max_batch=B), learning_rate=1e-3, weight_decay=0.01, beta=0.5).to(dev)
g = torch.Generator().manual_seed(0)
batch = ((0.365 * torch.randn(B, 1, 50, generator=g) + 0.019).clamp(-1, 1.3).to(dev),
         torch.log1p(0.0157 * torch.randn(B, 1, 100, generator=g).abs()).to(dev),
         torch.randint(1, 5, (B,), generator=g).to(dev))
for i in range(int(os.environ.get("WARM", "2"))):
    tm.training_step(batch, i)
    tm.optimizer.step(max_norm=1.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = tm.training_step(batch, 9)
tm.optimizer.step(max_norm=1.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
