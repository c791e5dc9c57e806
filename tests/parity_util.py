"""Shared helpers of the GPU parity tests: run the CUDA engine (through the C ABI) and the CPU oracle
on the same seeded inputs and report per-tensor errors."""
from __future__ import annotations

from collections import OrderedDict

import torch

from oracle import cvae_oracle as O


def make_engine(cfg: "O.CVAEConfig", max_batch: int, inference_only: bool = False, conv_path: int = 0):
    from hippie_b200.engine import Engine
    eng = Engine(cfg.z_dim, cfg.output_size_wave, cfg.output_size_isi, cfg.class_hidden_dim, cfg.num_sources,
                 cfg.num_classes, cfg.multimodal, max_batch, inference_only, conv_path)
    return eng.allocate("cuda:0")


def oracle_name(n: str) -> str:
    return n


def case_inputs(cfg, B, labelled, seed=1234, real_scale=False):
    x1, x2, labels, g = O.synthetic_batch(B, seed=seed, labelled=labelled)
    if cfg.output_size_wave != 50:
        x1 = torch.nn.functional.interpolate(x1, size=(cfg.output_size_wave,), mode="linear")
    if not cfg.multimodal:
        x2 = None
    if real_scale:  # large dynamic range like the real cellexplorer rows (index column leak)
        x1 = x1 * 40.0 + 100.0
    eps = torch.randn(B, cfg.z_dim, generator=g)
    if labelled:
        labels = labels.clone()
        labels[:, 0] = labels[:, 0] % cfg.num_classes
    return x1.contiguous(), (x2.contiguous() if x2 is not None else None), labels.contiguous(), eps.contiguous()


def perturbed_state(cfg, seed=42, jitter=0.1):
    """Reference init (bit-exact, oracle/cvae_oracle.py:init_state) with BatchNorm affine parameters and
    running statistics moved off their trivial 1/0 values so that every term is exercised."""
    st = O.init_state(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 7)
    for k in st:
        if k.endswith("running_mean"):
            st[k] = 0.05 * torch.randn(st[k].shape, generator=g)
        elif k.endswith("running_var"):
            st[k] = 1.0 + 0.2 * torch.rand(st[k].shape, generator=g)
        elif k.endswith("num_batches_tracked"):
            st[k] = torch.tensor(3)
        elif _is_bn_param(cfg, k):
            st[k] = st[k] + jitter * torch.randn(st[k].shape, generator=g)
    return st


_BN_CACHE = {}


def _is_bn_param(cfg, key):
    if cfg not in _BN_CACHE:
        names = set()
        for n, kind, _ in O.model_spec(cfg):
            if kind in ("ones", "zeros"):
                names.add(n)
        _BN_CACHE[cfg] = names
    return key in _BN_CACHE[cfg]


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def to_dtype(st, dtype):
    return OrderedDict((k, v.to(dtype) if v.is_floating_point() else v) for k, v in st.items())


def load_opt_state(eng, opt):
    """AdamW moments of the oracle (dict name -> tensor) into the engine's flat exp_avg / exp_avg_sq buffers."""
    for p in eng.params:
        if p.name in opt["exp_avg"]:
            eng.view_of(eng.exp_avg, p).copy_(opt["exp_avg"][p.name].to(torch.float32))
            eng.view_of(eng.exp_avg_sq, p).copy_(opt["exp_avg_sq"][p.name].to(torch.float32))


def run_train_case(cfg, B, labelled, seed=1234, beta=0.5, w1=1.0, w2=1.0, lr=1e-3, wd=0.01, clip=1.0,
                   real_scale=False, conv_path=0, engine=None, state=None, opt_state=None, inputs=None):
    """One teacher-forced optimisation step on the engine and on the oracle (fp32 and fp64).
    `state` / `opt_state` (fp32, oracle layout) start the step from a given point of a run -- parameters, BatchNorm
    buffers, AdamW moments and step counts -- instead of the perturbed initial state; `inputs` = (x1, x2, labels, eps)
    replaces the synthetic batch.  Returns a dict of comparisons."""
    dev = torch.device("cuda:0")
    x1, x2, labels, eps = inputs if inputs is not None else case_inputs(cfg, B, labelled, seed, real_scale)
    st = state if state is not None else perturbed_state(cfg)
    eng = engine or make_engine(cfg, max_batch=B, conv_path=conv_path)
    eng.load_named(st)
    if not eng.inference_only:
        eng.exp_avg.zero_(), eng.exp_avg_sq.zero_()
        if opt_state is not None:
            load_opt_state(eng, opt_state)
    cls, src = (labels.unbind(1) if labels.dim() == 2 else (None, labels))
    dx1, dx2 = x1.to(dev), (x2.to(dev) if x2 is not None else None)
    dsrc, dcls = src.contiguous().to(dev), (cls.contiguous().to(dev) if cls is not None else None)
    deps = eps.to(dev)
    scal, outs = eng.train_fwd_bwd(dx1, dx2, dsrc, dcls, deps, beta, w1, w2, outputs=True)
    torch.cuda.synchronize()
    launches_fb = eng.last_launch_count()
    grads = {k: v.detach().clone().cpu() for k, v in eng.named_grads().items()}
    taps_eng = {}
    tap_names = {t.name for t in eng.tensors}
    res = {"launches_fwd_bwd": launches_fb}

    # oracle, fp32 (the reference arithmetic) and fp64 (the yardstick)
    o32 = {}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        s = to_dtype(st, dt)
        opt = O.new_opt_state(s, cfg)
        if opt_state is not None:
            opt = {"step": dict(opt_state["step"]), "exp_avg": to_dtype(opt_state["exp_avg"], dt),
                   "exp_avg_sq": to_dtype(opt_state["exp_avg_sq"], dt)}
        new_st, new_opt, info = O.train_step(s, opt, cfg, x1.to(dt), x2.to(dt) if x2 is not None else None, labels,
                                             eps.to(dt), lr=lr, weight_decay=wd, beta=beta, w1=w1, w2=w2, max_norm=clip)
        with torch.no_grad():
            _, _, taps = O.forward(s, cfg, x1.to(dt), x2.to(dt) if x2 is not None else None, src, cls, eps.to(dt),
                                   train=True)
        o32[tag] = (new_st, new_opt, info, taps)
    new32, opt32, info32, taps32 = o32["f32"]
    new64, opt64, info64, taps64 = o32["f64"]

    # forward taps
    tap_err = {}
    for name, ref in taps64.items():
        if name in tap_names and ref.dim() == 3:
            got = eng.tensor_view(name, B).detach().cpu()
            scale = ref.abs().max().item() + 1e-30
            tap_err[name] = ((got.double() - ref).abs().max().item() / scale,
                             (taps32[name].double() - ref).abs().max().item() / scale)
    res["tap_err"] = tap_err
    # LeakyReLU sites whose sign differs from the fp64 oracle: each such element changes its local derivative by a
    # factor of 100, so gradients are only comparable tightly when there is none
    flips_eng = flips_f32 = 0
    for name, ref in taps64.items():
        if name in tap_names and ref.dim() == 3 and not name.endswith(".conv1") and not name.endswith(".linear"):
            got = eng.tensor_view(name, B).detach().cpu()
            flips_eng += int(((got > 0) != (ref > 0)).sum())
            flips_f32 += int(((taps32[name] > 0) != (ref > 0)).sum())
    res["flips_eng"], res["flips_f32"] = flips_eng, flips_f32
    out_err = {}
    for k in ("enc", "mu", "logvar", "dec1", "dec2"):
        if k in info64["out"]:
            ref = info64["out"][k]
            got = outs[k].detach().cpu().double().reshape(ref.shape)
            out_err[k] = ((got - ref).abs().max().item(), (info32["out"][k].double() - ref).abs().max().item(),
                          ref.abs().max().item())
    res["out_err"] = out_err
    s_eng = scal.detach().cpu().double()
    ref_l = torch.stack([info64["loss"], info64["mse1"], info64["mse2"], info64["kl"]])
    ref_l32 = torch.stack([info32["loss"], info32["mse1"], info32["mse2"], info32["kl"]]).double()
    res["loss_eng"], res["loss_f64"], res["loss_f32"] = s_eng[:4].tolist(), ref_l.tolist(), ref_l32.tolist()
    res["loss_rel"] = ((s_eng[:4] - ref_l).abs() / ref_l.abs().clamp_min(1e-12)).tolist()
    res["loss_rel_f32"] = ((ref_l32 - ref_l).abs() / ref_l.abs().clamp_min(1e-12)).tolist()

    # gradients
    g64, g32 = info64["grads_raw"], info32["grads_raw"]
    gnorm = torch.sqrt(sum((g.double() ** 2).sum() for g in g64.values())).item()
    gerr = {}
    flat_e, flat_r, flat_n = 0.0, 0.0, 0.0
    for n, ref in g64.items():
        e = (grads[n].double() - ref).norm().item()
        r = (g32[n].double() - ref).norm().item()
        gerr[n] = (e, r, ref.norm().item())
        res.setdefault("grad_dbg", {})[n] = (grads[n].double().norm().item(), (grads[n].double() * ref).sum().item() / (grads[n].double().norm().item() * ref.norm().item() + 1e-300))
        flat_e += e * e
        flat_r += r * r
        flat_n += ref.norm().item() ** 2
    res["grad_err"] = gerr
    # The same comparison with the LeakyReLU branches teacher-forced: the fp64 oracle is re-run with every backbone
    # activation taking the branch the ENGINE took (resp. the branch the reference's fp32 run took, for the yardstick),
    # so that inputs which rounding put on the other side of zero do not turn into factor-100 slope changes.
    sites = [n for n, ref in taps64.items()
             if n in tap_names and ref.dim() == 3 and not n.endswith(".conv1") and not n.endswith(".linear")]
    m_eng = {n: (eng.tensor_view(n, B).detach().cpu() > 0) for n in sites}
    m_f32 = {n: (taps32[n] > 0) for n in sites}
    s64 = to_dtype(st, torch.float64)

    def opt_of(dt):
        o = O.new_opt_state(to_dtype(st, dt), cfg)
        if opt_state is not None:
            o = {"step": dict(opt_state["step"]), "exp_avg": to_dtype(opt_state["exp_avg"], dt),
                 "exp_avg_sq": to_dtype(opt_state["exp_avg_sq"], dt)}
        return o
    x1d, x2d, epsd = x1.double(), (x2.double() if x2 is not None else None), eps.double()
    kw = dict(lr=lr, weight_decay=wd, beta=beta, w1=w1, w2=w2, max_norm=clip)
    _, _, i_me = O.train_step(s64, opt_of(torch.float64), cfg, x1d, x2d, labels, epsd, masks=m_eng, **kw)
    _, _, i_m32 = O.train_step(s64, opt_of(torch.float64), cfg, x1d, x2d, labels, epsd, masks=m_f32, **kw)
    gtf, fe, fr = {}, 0.0, 0.0
    for n, ref in g64.items():
        e = (grads[n].double() - i_me["grads_raw"][n]).norm().item()
        r = (g32[n].double() - i_m32["grads_raw"][n]).norm().item()
        gtf[n] = (e, r, ref.norm().item())
        fe += e * e
        fr += r * r
    res["grad_err_tf"] = gtf
    res["grad_flat_rel_tf"] = (fe ** 0.5) / (flat_n ** 0.5)
    res["grad_flat_rel_f32_tf"] = (fr ** 0.5) / (flat_n ** 0.5)
    res["loss_rel_tf"] = abs(float(s_eng[0]) - float(i_me["loss"])) / abs(float(i_me["loss"]))
    res["unforced_sites"] = [n for n, ref in taps64.items() if ref.dim() == 3 and n not in tap_names
                             and not n.endswith(".conv1") and not n.endswith(".linear")]
    res["grad_flat_rel"] = (flat_e ** 0.5) / (flat_n ** 0.5)
    res["grad_flat_rel_f32"] = (flat_r ** 0.5) / (flat_n ** 0.5)
    res["grad_global_norm"] = gnorm
    res["no_grad_params"] = [n for n in grads if n not in g64 and grads[n].abs().max().item() != 0.0]

    # BatchNorm running statistics after the step
    ns = {k: v.detach().clone().cpu() for k, v in eng.named_state().items()}
    run_err = 0.0
    for k, v in new64.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            run_err = max(run_err, (ns[k].double() - v).abs().max().item() / (v.abs().max().item() + 1e-6))
        if k.endswith("num_batches_tracked"):
            assert int(ns[k]) == int(v), k
    res["running_err"] = run_err

    # optimiser step
    has_cls = cls is not None
    step_no = 1 + (max(opt_state["step"].values()) if opt_state is not None else 0)
    step_cls = 1 + (opt_state["step"]["class_embedding.weight"] if opt_state is not None else 0)
    sc2 = eng.clip_adamw(lr, wd, step=step_no, max_norm=clip, step_cls=step_cls, has_cls_grad=has_cls)
    torch.cuda.synchronize()
    res["launches_opt"] = eng.last_launch_count()
    sc2 = sc2.cpu()
    res["grad_norm_eng"], res["grad_norm_f64"] = sc2[4].item(), info64["grad_norm"].item()
    res["clip_eng"], res["clip_f64"] = sc2[5].item(), info64["clip_coef"].item()
    ns2 = {k: v.detach().clone().cpu() for k, v in eng.named_state().items()}
    perr = 0.0
    for n in O.param_names(cfg):
        perr = max(perr, (ns2[n].double() - new64[n]).abs().max().item())
    res["param_abs_err"] = perr
    m_eng = {p.name: eng.view_of(eng.exp_avg, p).detach().cpu() for p in eng.params}
    me, mr = 0.0, 0.0
    for n in g64:
        me += (m_eng[n].double() - opt64["exp_avg"][n]).norm().item() ** 2
        mr += opt64["exp_avg"][n].norm().item() ** 2
    res["exp_avg_rel"] = (me / max(mr, 1e-300)) ** 0.5
    if not has_cls:
        res["cls_emb_untouched"] = bool(torch.equal(ns2["class_embedding.weight"], st["class_embedding.weight"]))
    return res, eng


def run_eval_case(cfg, B, labelled, seed=99, conv_path=0, engine=None, real_scale=False):
    dev = torch.device("cuda:0")
    x1, x2, labels, eps = case_inputs(cfg, B, labelled, seed, real_scale)
    st = perturbed_state(cfg)
    eng = engine or make_engine(cfg, max_batch=B, inference_only=True, conv_path=conv_path)
    eng.load_named(st)
    cls, src = (labels.unbind(1) if labels.dim() == 2 else (None, labels))
    dx1, dx2 = x1.to(dev), (x2.to(dev) if x2 is not None else None)
    dsrc, dcls = src.contiguous().to(dev), (cls.contiguous().to(dev) if cls is not None else None)
    scal = torch.zeros(8, device=dev)
    outs = eng.eval_forward(dx1, dx2, dsrc, dcls, eps.to(dev), 0.5, 1.0, 1.0, scalars=scal)
    emb = eng.embed(dx1, dx2, dsrc, dcls)
    emb_z = eng.embed(dx1, dx2, dsrc, dcls, zscore_ddof=0)
    torch.cuda.synchronize()
    s64 = to_dtype(st, torch.float64)
    with torch.no_grad():
        o64, _, _ = O.forward(s64, cfg, x1.double(), x2.double() if x2 is not None else None, src, cls, eps.double(),
                              train=False)
        tot, m1, m2, kl = O.loss_terms(o64, x1.double(), x2.double() if x2 is not None else None, 0.5, 1.0, 1.0,
                                       cfg.multimodal)
    res = {"abs_err": {}, "emb_err": {}}
    for k, ref in o64.items():
        res["abs_err"][k] = (outs[k].detach().cpu().double().reshape(ref.shape) - ref).abs().max().item()
    for k in ("enc", "mu", "logvar"):
        res["emb_err"][k] = (emb[k].detach().cpu().double() - o64[k]).abs().max().item()
    res["zscore_err"] = (emb_z["enc"].detach().cpu().double() - O.zscore_rows(o64["enc"], 0)).abs().max().item()
    emb_z1 = eng.embed(dx1, dx2, dsrc, dcls, zscore_ddof=1)  # pandas' default std (inference CLI, scripts/utils.py)
    res["zscore1_err"] = (emb_z1["enc"].detach().cpu().double() - O.zscore_rows(o64["enc"], 1)).abs().max().item()
    ref_l = torch.stack([tot, m1, m2, kl])
    res["loss_rel"] = ((scal[:4].cpu().double() - ref_l).abs() / ref_l.abs().clamp_min(1e-12)).tolist()
    return res, eng
