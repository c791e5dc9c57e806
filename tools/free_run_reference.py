"""Free-running loss curves of the REFERENCE's own classes at BASELINE.json's config 1 / 2 (multimodal pretrain, bs512,
z_dim 10, beta 0.5, AdamW lr 1e-3 wd 0.01, clip 1.0), frozen as tests/golden/free_run_bs512.npz.

SURVEY.md section 8c ("free-running envelope"): the reference is numerically chaotic after step 0 (F3), so a 1e-5 bound
is only meaningful teacher-forced; a free-running curve has to lie inside the envelope the reference spans with ITSELF
when only the rounding changes: {fp32 one thread, fp32 N threads, fp64}.  This script produces that envelope on the CPU
(build container: needs /root/reference or baseline/_ref; ~1 h on 8 cores for 200 steps) and the GPU test
(tests/test_gpu_parity.py::test_free_running_epoch_inside_reference_envelope) / tools/free_run_report.py run the engine
on the same units, eps and initial state.

Loop order = SURVEY.md section 3.2 (what Lightning does per step): training_step -> zero_grad -> backward ->
clip_grad_norm_(1.0) -> AdamW.step.  Units: oracle.synthetic_batch(512 * steps, seed 4242) in order, eps of step i =
torch.manual_seed(7000 + i); randn(512, z)  (what `randn_like` draws under that seed, hippie/model.py:399).

    python tools/free_run_reference.py [steps] [threads]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cvae_oracle as O  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

B, Z, SEED_DATA, SEED_EPS = 512, 10, 4242, 7000
HYPER = dict(lr=1e-3, wd=0.01, beta=0.5, clip=1.0)


def run(dtype, threads, steps, x1, x2, labels):
    R = load_reference()
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    base = R.model.MultiModalCVAE(Z, 50, 100, 5, 5, 5).to(dtype)
    mod = R.model.MultiModalCVAETrainModule(base, learning_rate=HYPER["lr"], weight_decay=HYPER["wd"], beta=HYPER["beta"])
    mod.train()
    losses = np.zeros((steps, 4))
    t0 = time.time()
    for i in range(steps):
        sl = slice(i * B, (i + 1) * B)
        batch = (x1[sl].to(dtype), x2[sl].to(dtype), labels[sl])
        torch.manual_seed(SEED_EPS + i)
        if dtype == torch.float64:  # same fp32 draw, cast (the fp64 generator stream differs)
            eps = torch.randn(B, Z).double()
            base.reparameterize = (lambda mu, lv, e=eps: mu + e * torch.exp(0.5 * lv))
            torch.manual_seed(SEED_EPS + i)
        loss = mod.training_step(batch, i)
        mod.optimizer.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(mod.parameters(), HYPER["clip"])
        mod.optimizer.step()
        lg = mod.logged
        losses[i] = [float(torch.as_tensor(lg[k]).detach()) for k in ("train_loss", "train_mse_loss1", "train_mse_loss2", "train_kl_loss")]
        if i % 20 == 0:
            print(f"[{dtype} t{threads}] step {i}: loss {losses[i, 0]:.6f}  ({time.time() - t0:.0f} s)", flush=True)
    return losses


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    x1, x2, labels, _ = O.synthetic_batch(B * steps, seed=SEED_DATA)
    out = {"steps": np.array(steps), "B": np.array(B), "z": np.array(Z), "seed_data": np.array(SEED_DATA),
           "seed_eps": np.array(SEED_EPS), "hyper": np.array([HYPER[k] for k in ("lr", "wd", "beta", "clip")]),
           "threads": np.array(threads)}
    out["f32_tN"] = run(torch.float32, threads, steps, x1, x2, labels)
    out["f64"] = run(torch.float64, threads, steps, x1, x2, labels)
    out["f32_t1"] = run(torch.float32, 1, steps, x1, x2, labels)
    path = os.path.join(ROOT, "tests", "golden", "free_run_bs512.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
