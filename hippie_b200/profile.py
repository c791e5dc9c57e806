"""Per-kernel-class timing of the implicit-GEMM convolutions for bench.py's roofline (CUDA events recorded by
the engine on the launching stream; see hippie_profile in include/hippie_b200.h)."""
from __future__ import annotations

import ctypes as C

import torch

KINDS = {0: "conv_fwd", 1: "conv_dgrad", 2: "conv_wgrad"}


def conv_roofline(eng, x1, x2, src, eps, beta=0.5):
    L, h = eng._L, eng._h
    scal = torch.zeros(8, device=eng.device)
    eng.train_fwd_bwd(x1, x2, src, None, eps, beta, 1.0, 1.0, scalars=scal)  # warm
    torch.cuda.synchronize()
    eng._check(L.hippie_profile(h, 1))
    eng.train_fwd_bwd(x1, x2, src, None, eps, beta, 1.0, 1.0, scalars=scal)
    torch.cuda.synchronize()
    out = {}
    tot_ms = tot_fl = 0.0
    for k, name in KINDS.items():
        ms, fl, n = C.c_double(), C.c_double(), C.c_int()
        eng._check(L.hippie_profile_read(h, k, C.byref(ms), C.byref(fl), C.byref(n)))
        if n.value:
            out[name] = {"launches": n.value, "total_ms": ms.value, "avg_us": 1e3 * ms.value / n.value,
                         "gflop": fl.value / 1e9, "tflops": fl.value / (ms.value * 1e-3) / 1e12 if ms.value > 0 else None}
            tot_ms += ms.value
            tot_fl += fl.value
    eng._check(L.hippie_profile(h, 0))
    out["all_conv"] = {"total_ms": tot_ms, "gflop": tot_fl / 1e9, "tflops": tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms else None}
    return out
