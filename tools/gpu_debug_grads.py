"""Diagnostic: activation-gradient tensors of the engine vs autograd of the oracle (fp64 and fp32)."""
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from oracle import cvae_oracle as O
import parity_util as U


def oracle_tap_grads(cfg, st, x1, x2, labels, eps, dt, beta=0.5):
    names = set(O.param_names(cfg))
    work = OrderedDict((k, (v.to(dt).clone().requires_grad_(True) if k in names else (v.to(dt) if v.is_floating_point() else v)))
                       for k, v in st.items())
    cls, src = (labels.unbind(1) if labels.dim() == 2 else (None, labels))
    out, _, taps = O.forward(work, cfg, x1.to(dt), x2.to(dt), src, cls, eps.to(dt), train=True)
    for t in taps.values():
        if t.requires_grad:
            t.retain_grad()
    total, *_ = O.loss_terms(out, x1.to(dt), x2.to(dt), beta, 1.0, 1.0, cfg.multimodal)
    total.backward()
    return {k: t.grad for k, t in taps.items() if t.grad is not None}, taps


def main():
    cfg = O.CVAEConfig(z_dim=int(os.environ.get("Z", "10")))
    B = int(os.environ.get("B", "48"))
    x1, x2, labels, eps = U.case_inputs(cfg, B, False)
    st = U.perturbed_state(cfg)
    eng = U.make_engine(cfg, B)
    eng.load_named(st)
    dev = torch.device("cuda:0")
    eng.train_fwd_bwd(x1.to(dev), x2.to(dev), labels.to(dev), None, eps.to(dev), 0.5, 1.0, 1.0)
    torch.cuda.synchronize()
    g64, t64 = oracle_tap_grads(cfg, st, x1, x2, labels, eps, torch.float64)
    g32, t32 = oracle_tap_grads(cfg, st, x1, x2, labels, eps, torch.float32)
    names = {t.name for t in eng.tensors}
    print("%-45s %10s %10s %10s" % ("tensor", "eng relL2", "f32 relL2", "norm"))
    for k, ref in g64.items():
        if ref.dim() != 3:
            continue
        for pref in ("g:", "d:"):
            if pref + k in names:
                got = eng.tensor_view(pref + k, B).detach().cpu().double()
                if got.shape != ref.shape:  # dilated gradient of a stride-2 conv
                    got = got[:, :, ::2][:, :, :ref.shape[2]]
                e = (got - ref).norm().item() / (ref.norm().item() + 1e-300)
                r = (g32[k].double() - ref).norm().item() / (ref.norm().item() + 1e-300)
                print("%-45s %10.3e %10.3e %10.3e" % (pref + k, e, r, ref.norm().item()))
    # LeakyReLU mask flips: elements whose sign differs between the engine and the fp64 oracle
    print("---- sign mismatches (engine vs f64 / f32 vs f64): count, max |ref| at a mismatch")
    tot_e = tot_r = 0
    for k, ref in t64.items():
        if ref.dim() == 3 and k in names and not k.endswith("conv1") and not k.endswith("linear"):
            got = eng.tensor_view(k, B).detach().cpu().double()
            mm = (got > 0) != (ref.detach() > 0)
            mr = (t32[k].detach() > 0) != (ref.detach() > 0)
            tot_e += int(mm.sum()); tot_r += int(mr.sum())
            if mm.any() or mr.any():
                print("%-40s eng %3d (max|ref| %.2e)   f32 %3d (max|ref| %.2e)  of %d" % (
                    k, int(mm.sum()), ref.detach().abs()[mm].max().item() if mm.any() else 0.0, int(mr.sum()),
                    ref.detach().abs()[mr].max().item() if mr.any() else 0.0, ref.numel()))
    print("total sign mismatches: engine %d, oracle-f32 %d" % (tot_e, tot_r))
    # forward taps, relative L2
    print("---- forward taps rel L2 (eng, f32)")
    for k, ref in t64.items():
        if ref.dim() == 3 and k in names:
            got = eng.tensor_view(k, B).detach().cpu().double()
            print("%-45s %10.3e %10.3e" % (k, (got - ref.detach()).norm().item() / ref.norm().item(),
                                           (t32[k].detach().double() - ref.detach()).norm().item() / ref.norm().item()))


if __name__ == "__main__":
    main()
