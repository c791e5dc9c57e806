"""Two-GPU tests (torchrun, NCCL; skipped on a one-GPU box -- run with `gpurun --gpus 2`): the data-parallel train step
equals the single-process computation on the concatenated batch with per-shard BatchNorm statistics (and the oracle's
step with the mean of the shard gradients), and the sharded inference CLI writes the CSV a single GPU writes."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")

_DP_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist
from oracle import cvae_oracle as O
import parity_util as U
from hippie_b200.parallel import train_step_overlapped, broadcast_state
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
cfg = O.CVAEConfig(z_dim=10)
PB = 40                                        # rows per rank; global batch 80
x1, x2, labels, eps = U.case_inputs(cfg, PB * world, False, seed=77)
st = U.perturbed_state(cfg)
from hippie_b200.engine import Engine
eng = Engine(cfg.z_dim, 50, 100, 5, cfg.num_sources, cfg.num_classes, True, PB).allocate(dev)
eng.load_named(st)
broadcast_state([eng.flat_params, eng.bn_mean, eng.bn_var])
sl = slice(rank * PB, (rank + 1) * PB)
scal = torch.zeros(8, device=dev)
args = (x1[sl].to(dev), x2[sl].to(dev), labels[sl].to(dev), None, eps[sl].to(dev), 0.5, 1.0, 1.0)
hist = []
for it in range(5):  # eager call, graph capture, graph replays: the exchange must wait for THIS step's slices every time
    eng.load_named(st)
    eng.exp_avg.zero_(), eng.exp_avg_sq.zero_()
    scale = train_step_overlapped(eng, *args, scalars=scal)
    torch.cuda.synchronize()
    g_dp = eng.flat_grads.clone()
    sc = eng.clip_adamw(1e-3, 0.01, 1, max_norm=1.0, grad_scale=scale)
    p_dp = eng.flat_params.clone()
    hist.append(g_dp)
    # every rank holds the same gradients and parameters after the exchange
    for t in (g_dp, p_dp):
        other = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(other, t)
        assert all(torch.equal(o, other[0]) for o in other), "ranks diverged"
for h in hist[1:]:
    assert ((h - hist[0]).norm() / hist[0].norm()).item() <= 1e-5, "replayed steps differ from the first"
if rank == 0:
    # single process: the shards one after the other on one engine, gradients summed
    ref = Engine(cfg.z_dim, 50, 100, 5, cfg.num_sources, cfg.num_classes, True, PB).allocate(dev)
    total = torch.zeros_like(g_dp)
    for r in range(world):
        ref.load_named(st)
        s = slice(r * PB, (r + 1) * PB)
        ref.train_fwd_bwd(x1[s].to(dev), x2[s].to(dev), labels[s].to(dev), None, eps[s].to(dev), 0.5, 1.0, 1.0)
        total += ref.flat_grads
    rel = ((g_dp - total).norm() / total.norm()).item()
    assert rel <= 1e-5, rel                       # same kernels; only the order of the fp32 atomics differs
    # the oracle: Lightning/DDP semantics = per-shard BatchNorm, mean of the shard gradients, then clip + AdamW
    s64 = U.to_dtype(st, torch.float64)
    shard = []
    for r in range(world):
        s = slice(r * PB, (r + 1) * PB)
        _, _, info = O.train_step(s64, O.new_opt_state(s64, cfg), cfg, x1[s].double(), x2[s].double(), labels[s], eps[s].double(),
                                  lr=1e-3, weight_decay=0.01, beta=0.5, max_norm=1.0)
        shard.append(info["grads_raw"])
    mean = {k: sum(g[k] for g in shard) / world for k in shard[0]}
    new64, _, info = O.train_step(s64, O.new_opt_state(s64, cfg), cfg, x1[:PB].double(), x2[:PB].double(), labels[:PB],
                                  eps[:PB].double(), lr=1e-3, weight_decay=0.01, beta=0.5, max_norm=1.0,
                                  grad_hook=lambda g: mean)
    got = {p.name: eng.view_of(g_dp, p).cpu().double() * scale for p in eng.params}
    num = sum(((got[k] - v) ** 2).sum() for k, v in mean.items()) ** 0.5
    den = sum((v ** 2).sum() for v in mean.values()) ** 0.5
    assert float(num / den) <= 0.03, float(num / den)   # free-running LeakyReLU branches: the gross-error bound
    assert abs(sc[4].item() - info["grad_norm"].item()) <= 5e-3 * info["grad_norm"].item()
    perr = max((eng.view_of(p_dp, p).cpu().double() - new64[p.name]).abs().max().item() for p in eng.params
               if p.name in mean)
    assert perr <= 2e-3 + 1e-6, perr
    print("dp ok: grads vs single process rel %.1e, vs oracle rel %.1e, param abs %.1e" % (rel, float(num / den), perr))
dist.barrier()
dist.destroy_process_group()
'''


def _torchrun(args, port, timeout=600):
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                           "127.0.0.1", "--master-port", str(port)] + args, capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


@needs2
def test_data_parallel_step_equals_single_process_and_oracle(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER)
    out = _torchrun([str(script), ROOT], 29581)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "dp ok" in out.stdout


@needs2
def test_sharded_inference_cli_writes_the_single_gpu_csv(tmp_path):
    """BASELINE.json configs[3] as a product path: `torchrun ... scripts/inference_from_trained_model.py` shards the units
    over the ranks (no communication until the final gather) and rank 0 writes the CSV; it must equal the one-GPU CSV
    byte for byte (eval-mode BatchNorm: units are independent)."""
    import pandas as pd
    from test_cli_gpu import _write_table
    from hippie_b200 import model as M
    root = str(tmp_path / "datasets")
    rng = np.random.default_rng(3)
    _write_table(root, "cellexplorer-celltype", 301, 46, 100, rng)  # odd count: ragged shards
    torch.manual_seed(1)
    m = M.MultiModalCVAETrainModule(M.MultiModalCVAE(10, 50, 100, 5, 5, 1, max_batch=64))
    ckpt = str(tmp_path / "joint.ckpt")
    torch.save({"state_dict": m.state_dict()}, ckpt)  # the part of a Lightning .ckpt the CLI reads
    cli = os.path.join(ROOT, "scripts", "inference_from_trained_model.py")
    common = ["--z_dim", "10", "--dataset", "cellexplorer-celltype", "--joint-checkpoint", ckpt, "--data-root", root, "--no-umap"]
    one = subprocess.run([sys.executable, cli] + common + ["--output-dir", str(tmp_path / "one")], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    two = _torchrun([cli] + common + ["--output-dir", str(tmp_path / "two")], 29583)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-2000:]
    name = "cellexplorer-celltype_joint_embeddings.csv"
    a, b = open(tmp_path / "one" / name, "rb").read(), open(tmp_path / "two" / name, "rb").read()
    assert len(pd.read_csv(tmp_path / "one" / name)) == 301
    assert a == b
