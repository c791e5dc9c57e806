"""Free-running loss curve of the engine at BASELINE.json's config 1 / 2 (multimodal pretrain, bs512, z 10, beta 0.5, AdamW
lr 1e-3 wd 0.01, clip 1.0) against the envelope the REFERENCE spans with itself when only the rounding changes
(tests/golden/free_run_bs512.npz: the reference's own classes in fp32 on N threads, fp32 on one thread and fp64, made by
tools/free_run_reference.py).  Same units, eps draws and seed-42 initial state.  SURVEY.md section 8c "free-running
envelope": the reference is chaotic after step 0, so the engine is compared with the spread between the reference's runs.

    python tools/free_run_report.py [out.md]        (on a B200)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cvae_oracle as O  # noqa: E402  (only its synthetic_batch: the units the reference curves were made on)


def engine_curve(fx):
    from hippie_b200 import model as M
    steps, B, Z = int(fx["steps"]), int(fx["B"]), int(fx["z"])
    lr, wd, beta, clip = [float(v) for v in fx["hyper"]]
    dev = torch.device("cuda:0")
    x1, x2, labels, _ = O.synthetic_batch(B * steps, seed=int(fx["seed_data"]))
    torch.manual_seed(42)
    m = M.MultiModalCVAE(Z, 50, 100, 5, 5, 5, max_batch=B)
    tm = M.MultiModalCVAETrainModule(m, learning_rate=lr, weight_decay=wd, beta=beta).to(dev)
    m.train()
    out = np.zeros((steps, 4))
    for i in range(steps):
        sl = slice(i * B, (i + 1) * B)
        torch.manual_seed(int(fx["seed_eps"]) + i)
        eps = torch.randn(B, Z).to(dev)
        tm.training_step((x1[sl], x2[sl], labels[sl]), i, eps=eps)
        tm.optimizer.step(max_norm=clip)
        out[i] = [float(tm.logged[k]) for k in ("train_loss", "train_mse_loss1", "train_mse_loss2", "train_kl_loss")]
    return out


def envelope(fx):
    """Per step: the largest relative distance between two of the reference's own runs, and its running maximum over a
    window of +-5 steps (a divergence shows up a few steps apart in different runs)."""
    runs = [fx[k][:, 0] for k in ("f32_tN", "f32_t1", "f64")]
    ref = fx["f64"][:, 0]
    spread = np.zeros(len(ref))
    for a in range(3):
        for b in range(a + 1, 3):
            spread = np.maximum(spread, np.abs(runs[a] - runs[b]) / np.abs(ref))
    win = np.array([spread[max(0, i - 5):i + 6].max() for i in range(len(ref))])
    return spread, win


def main():
    fx = np.load(os.path.join(ROOT, "tests", "golden", "free_run_bs512.npz"))
    got = engine_curve(fx)
    ref = fx["f64"][:, 0]
    spread, win = envelope(fx)
    dev = np.abs(got[:, 0] - ref) / np.abs(ref)
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_free_run_bs512.md")
    steps = len(ref)
    with open(out, "w") as f:
        f.write("# Free-running pretrain loss curve, bs512, z=10, beta=0.5 (BASELINE.json configs 1 / 2), %d steps\n\n" % steps)
        f.write("Reference = its own classes on the CPU (tests/golden/free_run_bs512.npz, tools/free_run_reference.py); engine = "
                "`MultiModalCVAETrainModule.training_step` + `FusedAdamW.step` on one B200, same units / eps / seed-42 init.\n"
                "`spread` = largest relative distance between two of the reference's own runs {fp32 N threads, fp32 1 thread, "
                "fp64} at that step; `dev` = |engine - reference fp64| / reference fp64.\n\n")
        f.write("| step | ref fp64 | ref fp32 (N thr) | ref fp32 (1 thr) | engine | spread | dev |\n|---:|---:|---:|---:|---:|---:|---:|\n")
        for i in sorted(set([0, 1, 2, 3, 5, 8, 10, 15, 20, 30, 50, 75, 100, 125, 150, 175, steps - 1])):
            if i < steps:
                f.write(f"| {i} | {ref[i]:.6f} | {fx['f32_tN'][i, 0]:.6f} | {fx['f32_t1'][i, 0]:.6f} | {got[i, 0]:.6f} | {spread[i]:.1e} | {dev[i]:.1e} |\n")
        f.write(f"\nstep 0 (identical state): dev {dev[0]:.1e} (bound 1e-5).  max over steps: spread {spread.max():.2e}, dev {dev.max():.2e}; "
                f"steps with dev > windowed spread: {(dev > win).sum()} of {steps}; with dev > 4 x windowed spread + 2e-3: "
                f"{(dev > 4 * win + 2e-3).sum()}.\nmean of the last 20 steps: reference fp64 {ref[-20:].mean():.6f}, engine "
                f"{got[-20:, 0].mean():.6f} (rel {abs(got[-20:, 0].mean() - ref[-20:].mean()) / ref[-20:].mean():.1e}; reference "
                f"fp32 N-thread rel {abs(fx['f32_tN'][-20:, 0].mean() - ref[-20:].mean()) / ref[-20:].mean():.1e}).\n")
    np.savez_compressed(os.path.splitext(out)[0] + ".npz", engine=got, dev=dev, spread=spread)
    print(open(out).read())


if __name__ == "__main__":
    main()
