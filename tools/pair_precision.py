"""CPU emulation of the operand formats the tcgen05 conv path can use, to pick one BEFORE writing kernels.

For every Conv1d of the forward pass (stem excluded: it runs on CUDA cores in fp32) the operands are rounded the way a
given tensor-core scheme would see them, the contraction itself is done in fp64 (so only the operand representation
error is measured) and the result is rounded to fp32.  Reported: loss / embedding error against the fp64 oracle,
beside the plain fp32 oracle's own error.  Test infrastructure only (imports oracle/).

  python tools/pair_precision.py            # real cellexplorer rows (dynamic range up to 391) + synthetic bs64
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F

from oracle import cvae_oracle as O

W_SCALE = 2.0 ** 8


def split16(x, dt):
    h0 = x.to(dt)
    h1 = (x - h0.to(x.dtype)).to(dt)
    return h0.double(), h1.double()


def tf32_rn(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def make_conv(scheme):
    def conv(cx, name, x, stride, padding):
        w = cx.st[name + ".weight"]
        b = cx.st.get(name + ".bias")
        if x.dtype != torch.float32 or w.shape[1] == 1 or scheme == "fp32":
            return F.conv1d(x, w, b, stride=stride, padding=padding)
        if scheme == "fp16x2":  # x = h0 + h1 (fp16), w * 2^8 = w0 + w1 (fp16); h0w0 + h0w1 + h1w0
            a0, a1 = split16(x, torch.float16)
            w0, w1 = split16(w * W_SCALE, torch.float16)
            y = (F.conv1d(a0, w0, None, stride, padding) + F.conv1d(a0, w1, None, stride, padding) +
                 F.conv1d(a1, w0, None, stride, padding)) / W_SCALE
        elif scheme == "bf16x2":
            a0, a1 = split16(x, torch.bfloat16)
            w0, w1 = split16(w, torch.bfloat16)
            y = (F.conv1d(a0, w0, None, stride, padding) + F.conv1d(a0, w1, None, stride, padding) +
                 F.conv1d(a1, w0, None, stride, padding))
        elif scheme == "bf16x3":
            a0, a1 = split16(x, torch.bfloat16)
            a2 = (x.double() - a0 - a1).to(torch.bfloat16).double()
            w0, w1 = split16(w, torch.bfloat16)
            w2 = (w.double() - w0 - w1).to(torch.bfloat16).double()
            y = sum(F.conv1d(p, q, None, stride, padding) for p, q in
                    [(a0, w0), (a0, w1), (a1, w0), (a1, w1), (a0, w2), (a2, w0)])
        elif scheme == "tf32x3":
            a0 = tf32_rn(x)
            a1 = x - a0
            w0 = tf32_rn(w)
            w1 = w - w0
            y = (F.conv1d(a0.double(), w0.double(), None, stride, padding) +
                 F.conv1d(a0.double(), tf32_rn(w1).double(), None, stride, padding) +
                 F.conv1d(tf32_rn(a1).double(), w0.double(), None, stride, padding))
        elif scheme == "tf32x1":
            y = F.conv1d(tf32_rn(x).double(), tf32_rn(w).double(), None, stride, padding)
        elif scheme == "fp16x1":
            y = F.conv1d(x.half().double(), (w * W_SCALE).half().double(), None, stride, padding) / W_SCALE
        else:
            raise ValueError(scheme)
        y = y.float()
        return y if b is None else y + b.view(1, -1, 1)
    return conv


def run(cfg, st, x1, x2, src, eps, scheme, train):
    orig = O._conv1d
    O._conv1d = make_conv(scheme)
    try:
        out, _, _ = O.forward(st, cfg, x1, x2, src, None, eps, train=train)
        loss = O.loss_terms(out, x1, x2, 0.5)
    finally:
        O._conv1d = orig
    return out, [float(v) for v in loss]


def report(tag, cfg, st, x1, x2, src, eps):
    st64 = {k: (v.double() if v.is_floating_point() else v) for k, v in st.items()}
    for train in (True, False):
        ref, l64 = run(cfg, st64, x1.double(), x2.double(), src, eps.double(), "fp32", train)
        print(f"--- {tag}  train={train}  loss(fp64) = {l64[0]:.8f}")
        for scheme in ("fp32", "tf32x3", "fp16x2", "bf16x3", "bf16x2", "tf32x1", "fp16x1"):
            out, l = run(cfg, st, x1, x2, src, eps, scheme, train)
            lrel = max(abs(a - b) / abs(b) for a, b in zip(l, l64) if b != 0)
            emb = max(float((out[k].double() - ref[k]).abs().max()) for k in ("enc", "mu"))
            dec = max(float((out[k].double() - ref[k]).abs().max()) for k in ("dec1", "dec2"))
            print(f"  {scheme:8s} loss rel {lrel:.2e}   enc/mu abs {emb:.2e}   dec abs {dec:.2e}")


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.CVAEConfig(z_dim=10)
    st = O.init_state(cfg, seed=42)
    z = np.load(os.path.join(ROOT, "tests", "golden", "cellexplorer_raw48.npz"))
    x1, x2 = torch.from_numpy(z["x1"]), torch.from_numpy(z["x2"])
    src = torch.full((x1.shape[0],), 3, dtype=torch.int64)
    eps = torch.randn(x1.shape[0], cfg.z_dim, generator=torch.Generator().manual_seed(5))
    print("x1 range", float(x1.min()), float(x1.max()))
    report("real cellexplorer rows, B=48", cfg, st, x1, x2, src, eps)
    x1, x2, labels, g = O.synthetic_batch(128, seed=1234, labelled=False)
    eps = torch.randn(128, cfg.z_dim, generator=g)
    # a trained-like state: perturb the BatchNorm affine parameters and running statistics
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_util as U
    report("synthetic, B=128, perturbed BN", cfg, U.perturbed_state(cfg), x1, x2, labels, eps)


if __name__ == "__main__":
    main()
