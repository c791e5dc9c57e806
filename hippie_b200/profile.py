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


def graph_replay_kernel_times(eng, x1, x2, src, eps, beta=0.5, pattern="conv_pair_kernel"):
    """CUPTI durations (torch.profiler) of the kernels matching `pattern` inside ONE graph-replayed train step -- the
    step as bench.py times it, two branch chains and the weight-gradient streams running concurrently -- next to the
    event-bracketed eager figures of conv_roofline.  Returns {"launches", "total_ms", "avg_us"} or {"error"}."""
    import json
    import os
    import tempfile

    from torch.profiler import ProfilerActivity, profile
    scal = torch.zeros(8, device=eng.device)
    for _ in range(4):  # the first calls of a signature run eagerly, then the graph is captured
        eng.train_fwd_bwd(x1, x2, src, None, eps, beta, 1.0, 1.0, scalars=scal)
    torch.cuda.synchronize()
    steps = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            eng.train_fwd_bwd(x1, x2, src, None, eps, beta, 1.0, 1.0, scalars=scal)
        torch.cuda.synchronize()
    fd, path = tempfile.mkstemp(suffix=".json")
    os.close(fd)
    try:
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel" and pattern in e.get("name", "")]
    finally:
        os.unlink(path)
    if not ev or len(ev) % steps:
        return {"error": f"{len(ev)} matching kernel events over {steps} steps"}
    tot_us = sum(e["dur"] for e in ev) / steps
    n = len(ev) // steps
    return {"launches": n, "total_ms": tot_us / 1e3, "avg_us": tot_us / n}
