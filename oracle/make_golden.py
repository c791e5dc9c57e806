"""Freeze golden vectors from the REFERENCE's own classes (build container only).

Run:  python oracle/make_golden.py [--out DIR]   (needs /root/reference; writes tests/golden/ or DIR)

What it does
  1. imports /root/reference/hippie/{backbones,model,dataloading}.py unmodified (oracle/ref_loader.py: by path, so
     that this repository's `hippie/` alias package cannot shadow them; `pytorch_lightning` stand-in of oracle/_plstub);
  2. drives the reference `MultiModalCVAETrainModule` / `hippieUnimodalEmbeddingModelCVAE`
     with the Lightning call order of SURVEY.md §3.2 (training_step -> zero_grad -> backward
     -> clip_grad_norm_(1.0) -> AdamW.step) on seeded inputs;
  3. asserts that oracle/cvae_oracle.py reproduces every reference quantity (this is what
     pins the oracle), and
  4. stores compact summaries (inputs, outputs, losses, per-tensor gradient norms and
     leading elements, post-step parameter samples, init checksums) as tests/golden/*.npz.
The fixtures travel with the repo; /root/reference does not.
"""
import os
import sys

from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import cvae_oracle as O  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

torch.set_num_threads(8)
NHEAD = 6  # leading elements of every tensor kept in the fixture
OUT_DIR = os.path.join(ROOT, "tests", "golden")  # `python oracle/make_golden.py --out DIR` writes elsewhere (CI re-check)


def ref_modules():
    # loaded by path: this repository's own `hippie/` alias package would shadow the reference's namespace package
    R = load_reference("/root/reference")
    return R.model, R.dataloading


def heads(d, names):
    return np.stack([np.pad(d[n].detach().double().flatten()[:NHEAD].numpy(), (0, max(0, NHEAD - d[n].numel())))
                     for n in names])


def _build_ref(cfg, hyper, dtype):
    RM, _ = ref_modules()
    torch.manual_seed(42)
    if cfg.multimodal:
        base = RM.MultiModalCVAE(cfg.z_dim, cfg.output_size_wave, cfg.output_size_isi, cfg.class_hidden_dim,
                                 cfg.num_sources, cfg.num_classes)
    else:
        base = RM.hippieUnimodalCVAE(cfg.z_dim, cfg.output_size_wave, cfg.class_hidden_dim, cfg.num_sources,
                                     cfg.num_classes)
    base = base.to(dtype)
    if cfg.multimodal:
        mod = RM.MultiModalCVAETrainModule(base, learning_rate=hyper["lr"], weight_decay=hyper["wd"],
                                           beta=hyper["beta"], mod1_weight=hyper["w1"], mod2_weight=hyper["w2"])
    else:
        mod = RM.hippieUnimodalEmbeddingModelCVAE(base, learning_rate=hyper["lr"], weight_decay=hyper["wd"],
                                                  beta=hyper["beta"])
    return base, mod


def _eps(seed, b, z, dtype):
    torch.manual_seed(seed)
    return torch.randn(b, z).to(dtype)


def run_case(tag, cfg, x1, x2, labels, eps_seeds, hyper, steps):
    """Drives the reference in fp32 (unmodified: eps comes from seeding the default generator,
    as the reference draws it) and in fp64 (the "truth" yardstick; eps injected by replacing
    the instance's `reparameterize` with the same fp32 draw cast to fp64), and pins the oracle
    to both: fp64 to 1e-9 everywhere, fp32 within the reference's own fp32-vs-fp64 noise."""
    pnames = O.param_names(cfg)
    B = x1.shape[0]
    cls, src = (labels.unbind(1) if labels.dim() == 2 else (None, labels))
    keys = ["enc", "mu", "logvar", "dec1"] + (["dec2"] if cfg.multimodal else [])
    fx = {"param_names": np.array(pnames), "x1": x1.numpy(), "labels": labels.numpy(),
          "eps_seeds": np.array(eps_seeds),
          "hyper": np.array([hyper[k] for k in ("lr", "wd", "beta", "w1", "w2", "clip")], dtype=np.float64)}
    if x2 is not None:
        fx["x2"] = x2.numpy()
    for dtype, dt in ((torch.float32, "f32"), (torch.float64, "f64")):
        base, mod = _build_ref(cfg, hyper, dtype)
        st = O.init_state(cfg, seed=42)
        ref_sd = base.state_dict()
        assert list(ref_sd.keys()) == list(st.keys()), "state_dict order differs"
        assert pnames == [n for n, _ in base.named_parameters()]
        if dtype == torch.float32:
            for k in st:
                assert torch.equal(ref_sd[k], st[k]), f"init differs at {k}"
            fx["init_sum"] = np.array([st[n].double().sum().item() for n in pnames])
            fx["init_abs"] = np.array([st[n].double().abs().sum().item() for n in pnames])
            fx["init_head"] = heads(st, pnames)
        st = OrderedDict((k, (v.to(dtype) if v.is_floating_point() else v)) for k, v in st.items())
        X1 = x1.to(dtype)
        X2 = x2.to(dtype) if x2 is not None else None
        batch = (X1, X2, labels) if cfg.multimodal else (X1, labels)
        opt = O.new_opt_state(st, cfg)
        mod.train()
        for s in range(steps):
            eps = _eps(eps_seeds[s], B, cfg.z_dim, dtype)
            # ---- reference --------------------------------------------------------------
            if dtype == torch.float64:
                base.reparameterize = (lambda mu, lv, e=eps: mu + e * torch.exp(0.5 * lv))
            torch.manual_seed(eps_seeds[s])
            if s == 0:  # train-mode forward outputs at the initial weights (no state change wanted:
                # snapshot + restore the BN buffers the forward mutates)
                snap = {k: v.clone() for k, v in base.state_dict().items()}
                with torch.no_grad():
                    r0 = mod(batch)
                base.load_state_dict(snap)
                torch.manual_seed(eps_seeds[s])
            loss = mod.training_step(batch, s)
            mod.optimizer.zero_grad()
            loss.backward()
            rgrads = {n: p.grad.clone() for n, p in base.named_parameters() if p.grad is not None}
            rnorm = torch.nn.utils.clip_grad_norm_(mod.parameters(), hyper["clip"])
            mod.optimizer.step()
            rlog = {k: float(v.detach()) for k, v in mod.logged.items()}
            # ---- oracle -----------------------------------------------------------------
            if s == 0:
                with torch.no_grad():
                    o0, _, _ = O.forward(st, cfg, X1, X2, src, cls, eps, train=True)
                for k, r in zip(keys, r0):
                    tol = (1e-10 if dtype == torch.float64 else 2e-5) * max(1.0, r.abs().max().item())
                    assert (o0[k] - r).abs().max().item() <= tol, (tag, dt, "fwd", k, (o0[k] - r).abs().max())
                    fx[f"{dt}_fwd0_{k}"] = r.numpy()
            st, opt, info = O.train_step(st, opt, cfg, X1, X2, labels, eps, lr=hyper["lr"], weight_decay=hyper["wd"],
                                         beta=hyper["beta"], w1=hyper["w1"], w2=hyper["w2"], max_norm=hyper["clip"])
            # ---- pin ----------------------------------------------------------------------
            gn = [n for n in pnames if n in rgrads]
            assert set(rgrads) == set(info["grads_raw"])
            gflat_r = torch.cat([rgrads[n].flatten() for n in gn]).double()
            gflat_o = torch.cat([info["grads_raw"][n].flatten() for n in gn]).double()
            rel = ((gflat_r - gflat_o).norm() / gflat_r.norm()).item()
            lrel = abs(info["loss"].item() - loss.item()) / abs(loss.item())
            new_sd = base.state_dict()
            if dtype == torch.float64:
                assert lrel < 1e-11, (tag, dt, s, lrel)
                assert rel < 1e-9, (tag, dt, s, "flat grad rel-L2", rel)
                assert abs(rnorm.item() - info["grad_norm"].item()) <= 1e-10 * rnorm.item()
                for k in new_sd:
                    if new_sd[k].is_floating_point():
                        assert torch.allclose(new_sd[k], st[k], rtol=1e-8, atol=1e-11), (tag, dt, s, k)
                    else:
                        assert int(new_sd[k]) == int(st[k]) == s + 1
                fx[f"f64_s{s}_gflat_ref"] = np.array(0.0)
                g64 = gflat_r
            else:
                tol = 2e-6 if s == 0 else 5e-3  # chaos after the first update
                assert lrel <= tol, (tag, dt, s, lrel)
                if s == 0:
                    assert rel < 2e-2, (tag, dt, "flat grad rel-L2", rel)
                    for n in gn:
                        d = (new_sd[n] - st[n]).abs().max().item()
                        assert d <= 2.0 * hyper["lr"] + 1e-7, (tag, n, d)
            print(f"[{tag}/{dt}] step{s} pinned: loss {loss.item():.10f} (oracle rel {lrel:.1e}); "
                  f"flat-grad rel-L2 oracle vs reference {rel:.2e}; |g| {rnorm.item():.6f}")
            fx[f"{dt}_s{s}_loss"] = np.array(
                [rlog.get("train_loss"), rlog.get("train_mse_loss1", rlog.get("train_mse_loss")),
                 rlog.get("train_mse_loss2", 0.0), rlog.get("train_kl_loss")], dtype=np.float64)
            fx[f"{dt}_s{s}_grad_norm"] = np.array(rnorm.item())
            fx[f"s{s}_grad_names"] = np.array(gn)
            fx[f"{dt}_s{s}_grad_l2"] = np.array([rgrads[n].double().norm().item() for n in gn])
            fx[f"{dt}_s{s}_grad_head"] = heads(rgrads, gn)
            fx[f"{dt}_s{s}_param_head"] = heads(new_sd, pnames)
            fx[f"{dt}_s{s}_param_sum"] = np.array([new_sd[n].double().sum().item() for n in pnames])
            rn = [k for k in new_sd if "running" in k]
            fx["running_names"] = np.array(rn)
            fx[f"{dt}_s{s}_running_head"] = heads(new_sd, rn)
        # ---- eval-mode forward on the post-training weights ---------------------------------
        mod.eval()
        eps = _eps(777, B, cfg.z_dim, dtype)
        if dtype == torch.float64:
            base.reparameterize = (lambda mu, lv, e=eps: mu + e * torch.exp(0.5 * lv))
        torch.manual_seed(777)
        with torch.no_grad():
            routs = mod(batch)
            # teacher-forced: evaluate the oracle on the REFERENCE's weights
            ot, _, _ = O.forward(OrderedDict(base.state_dict()), cfg, X1, X2, src, cls, eps, train=False)
        for k, r in zip(keys, routs):
            tol = (1e-10 if dtype == torch.float64 else 2e-5) * max(1.0, r.abs().max().item())
            assert (ot[k] - r).abs().max().item() <= tol, (tag, dt, "eval", k, (ot[k] - r).abs().max())
            fx[f"{dt}_eval_{k}"] = r.numpy()
        print(f"[{tag}/{dt}] eval-mode forward pinned")
    fx["eval_eps_seed"] = np.array(777)
    out = os.path.join(OUT_DIR, f"{tag}.npz")
    np.savez_compressed(out, **fx)
    print(f"[{tag}] wrote {out} ({os.path.getsize(out) / 1024:.1f} KiB)")


def real_batch(n):
    """First n units of datasets/cellexplorer-celltype through the REFERENCE EphysDataset
    (mode='both', normalize=False) exactly as the training script reads them
    (pd.read_csv without index_col -> the index column is a feature)."""
    import pandas as pd
    _, RD = ref_modules()
    wf = pd.read_csv("/root/reference/datasets/cellexplorer-celltype/waveforms.csv").to_numpy()
    isi = pd.read_csv("/root/reference/datasets/cellexplorer-celltype/isi_dist.csv").to_numpy()
    ds = RD.EphysDataset(wf, isi, mode="both", normalize=False)
    xs = [ds[i] for i in range(n)]
    x1 = torch.stack([a for a, _ in xs])
    x2 = torch.stack([b for _, b in xs])
    # pin the oracle's dataset transform on every unit of the file
    for i in range(len(ds)):
        a, b = ds[i]
        oa, ob = O.dataset_item(wf[i], isi[i])
        assert torch.equal(a, oa) and torch.equal(b, ob), f"dataset transform differs at row {i}"
    print(f"[data] oracle dataset transform bit-exact on all {len(ds)} cellexplorer-celltype units")
    return wf[:n], isi[:n], x1, x2


def main():
    hyper = {"lr": 1e-3, "wd": 0.01, "beta": 0.5, "w1": 1.0, "w2": 1.0, "clip": 1.0}
    # A. real data, label-free pretrain-style batch (source id 3), z=10
    wf, isi, x1, x2 = real_batch(48)
    np.savez_compressed(os.path.join(OUT_DIR, "cellexplorer_raw48.npz"), wf=wf, isi=isi,
                        x1=x1.numpy(), x2=x2.numpy())
    run_case("real48_z10", O.CVAEConfig(z_dim=10), x1, x2, torch.full((48,), 3, dtype=torch.long), [101, 102],
             hyper, steps=2)
    # B. synthetic, labelled (supervised finetune shape, [class, source]), z=10, bs64
    s1, s2, lab, _ = O.synthetic_batch(64, seed=1234, labelled=True)
    run_case("synth64_labelled_z10", O.CVAEConfig(z_dim=10, num_classes=4), s1, s2, lab, [201, 202],
             dict(hyper, lr=1e-4), steps=2)
    # C. z sweep member, label-free, weights != 1
    s1, s2, lab, _ = O.synthetic_batch(24, seed=99, labelled=False)
    run_case("synth24_z32", O.CVAEConfig(z_dim=32), s1, s2, lab, [301],
             dict(hyper, w1=0.7, w2=1.3, beta=1.0), steps=1)
    # D. unimodal twins (wave L=50 and isi L=100), default --model-type of the CLI
    s1, s2, lab, _ = O.synthetic_batch(24, seed=7, labelled=True)
    run_case("uni_wave24_z10", O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=50), s1, None, lab, [401],
             dict(hyper, beta=1.0), steps=1)
    run_case("uni_isi24_z10", O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=100), s2, None, lab, [402],
             dict(hyper, beta=1.0), steps=1)


if __name__ == "__main__":
    if "--out" in sys.argv:
        OUT_DIR = sys.argv[sys.argv.index("--out") + 1]
        os.makedirs(OUT_DIR, exist_ok=True)
    main()
