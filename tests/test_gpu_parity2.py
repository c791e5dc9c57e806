"""Round-2 parity cases ("next round" 1): BASELINE.json's own sizes against the oracle (bs512 pretrain
step, bs64 supervised step for z in {10, 32, 64}), teacher-forced parity k steps into a run on the real cellexplorer
rows, the AdamW kernel fed the oracle's gradient, per-row z-score with ddof 1, the device-side saturation / label flags."""
import os

import numpy as np
import pytest
import torch

from oracle import cvae_oracle as O
import parity_util as U
from test_gpu_parity import LOSS_RTOL, EMB_ATOL, _check_train

pytestmark = pytest.mark.gpu


def test_bs512_benchmark_step_matches_oracle():
    """BASELINE.json configs[1]: the bs512 label-free pretrain step (z=10, beta 0.5) teacher-forced against the oracle:
    different weight-gradient split counts, BatchNorm chunking and head CTA ranges than the small cases."""
    res, eng = U.run_train_case(O.CVAEConfig(z_dim=10), 512, False)
    assert eng.conv_path_in_use() == 2
    _check_train(res)


@pytest.mark.parametrize("z", [10, 32, 64])
def test_supervised_bs64_z_sweep_matches_oracle(z):
    """BASELINE.json configs[4]: supervised finetune step, bs64 with class labels, z_dim sweep."""
    res, _ = U.run_train_case(O.CVAEConfig(z_dim=z, num_classes=4), 64, True, lr=1e-4)
    _check_train(res, lr=1e-4)


ULP = 2.0 ** -23  # spacing of fp32 numbers relative to their magnitude (an upper bound of one ulp)


@pytest.mark.parametrize("clip", [1e9, 1.0])
def test_adamw_kernel_matches_torch_order_to_ulps(clip):
    """`adamw_kernel` fed a given gradient and given moments (step 7 of a run): p, m, v against the oracle's restatement of
    torch.optim.AdamW's single-tensor update (oracle/cvae_oracle.py:adamw_update) evaluated in fp32.  With clip = 1e9 the
    coefficient is exactly 1; with clip = 1.0 the oracle applies the coefficient the device computed (its own sum order
    differs from torch's by rounding), so both runs isolate the element-wise update.  Bound: 2 ulp of the operands of
    each update (m = m + (g - m) * 0.1 cancels, so the result itself can be arbitrarily smaller than its rounding error)."""
    cfg = O.CVAEConfig(z_dim=10, num_classes=4)
    eng = U.make_engine(cfg, 8)
    st = U.perturbed_state(cfg)
    eng.load_named(st)
    g = torch.Generator().manual_seed(11)
    names = O.param_names(cfg)
    grads = {n: 0.05 * torch.randn(st[n].shape, generator=g) for n in names}
    m0 = {n: 0.01 * torch.randn(st[n].shape, generator=g) for n in names}
    v0 = {n: 1e-4 * torch.rand(st[n].shape, generator=g) for n in names}
    for p in eng.params:
        eng.view_of(eng.flat_grads, p).copy_(grads[p.name])
    U.load_opt_state(eng, {"exp_avg": m0, "exp_avg_sq": v0})
    step, lr, wd = 7, 1e-3, 0.01
    sc = eng.clip_adamw(lr, wd, step=step, max_norm=clip, step_cls=step, has_cls_grad=True).cpu()
    torch.cuda.synchronize()
    total, coef = O.clip_coef(grads, clip)
    assert abs(sc[4].item() - total.item()) <= 1e-5 * total.item()  # fp32 sum of 16 M squares, another order
    assert abs(sc[5].item() - coef.item()) <= 1e-5 * coef.item()
    coef_dev = sc[5].clone()  # fp32, what the kernel multiplied with
    if clip >= 1e8:
        assert coef_dev.item() == 1.0
    new = {k: v.detach().cpu() for k, v in eng.named_state().items()}
    worst = {"p": 0.0, "m": 0.0, "v": 0.0}
    for p in eng.params:
        n = p.name
        gc = grads[n] * coef_dev
        pr, mr, vr = O.adamw_update(st[n], gc, m0[n], v0[n], step, lr, wd)
        got_m, got_v = eng.view_of(eng.exp_avg, p).cpu(), eng.view_of(eng.exp_avg_sq, p).cpu()
        scale_m = torch.maximum(m0[n].abs(), 0.1 * gc.abs())
        scale_v = torch.maximum(v0[n], 0.001 * gc * gc)
        upd = (st[n] - pr).abs()  # decay + step: p's own rounding plus the rounding of the update term ...
        denom = vr.sqrt() / (1 - 0.999 ** step) ** 0.5 + 1e-8
        carry = (lr / (1 - 0.9 ** step)) / denom * scale_m  # ... which also carries m's (cancellation-scale) rounding
        worst["m"] = max(worst["m"], ((got_m - mr).abs() / (ULP * scale_m)).max().item())
        worst["v"] = max(worst["v"], ((got_v - vr).abs() / (ULP * scale_v)).max().item())
        worst["p"] = max(worst["p"], ((new[n] - pr).abs() / (ULP * (st[n].abs() + 4 * upd + 4 * carry))).max().item())
    assert worst["m"] <= 2 and worst["v"] <= 2 and worst["p"] <= 2, worst


@pytest.mark.parametrize("k", [10, 30])
def test_teacher_forced_parity_k_steps_into_a_run_on_real_rows(golden_dir, k):
    """on the real cellexplorer rows (leaked index column, values up to ~400) operand magnitudes grow
    over the first 10-30 steps.  The oracle advances k steps in fp64 on the 48 real rows; the engine then takes step
    k + 1 from that state (parameters, BatchNorm buffers, AdamW moments, step counts) and must meet the same bounds as
    at step 0."""
    fx = np.load(os.path.join(golden_dir, "real48_z10.npz"))
    cfg = O.CVAEConfig(z_dim=10)
    x1, x2, labels = torch.tensor(fx["x1"]), torch.tensor(fx["x2"]), torch.tensor(fx["labels"])
    st = U.to_dtype(O.init_state(cfg, seed=42), torch.float64)
    opt = O.new_opt_state(st, cfg)
    g = torch.Generator().manual_seed(500 + k)
    kw = dict(lr=1e-3, weight_decay=0.01, beta=0.5, max_norm=1.0)
    for _ in range(k):
        eps = torch.randn(48, cfg.z_dim, generator=g)
        st, opt, info = O.train_step(st, opt, cfg, x1.double(), x2.double(), labels, eps.double(), **kw)
    st32 = U.to_dtype(st, torch.float32)
    opt32 = {"step": dict(opt["step"]), "exp_avg": U.to_dtype(opt["exp_avg"], torch.float32),
             "exp_avg_sq": U.to_dtype(opt["exp_avg_sq"], torch.float32)}
    eps = torch.randn(48, cfg.z_dim, generator=g)
    res, _ = U.run_train_case(cfg, 48, False, state=st32, opt_state=opt32, inputs=(x1, x2, labels, eps))
    _check_train(res)


@pytest.mark.parametrize("name", ["mm", "uni"])
def test_zscore_ddof1_matches_oracle(name):
    """The inference CLI z-scores every embedding row with pandas' default std (ddof 1, reference
    scripts/inference_from_trained_model.py); the fused device version against oracle.zscore_rows(., 1)."""
    cfg = O.CVAEConfig(z_dim=10) if name == "mm" else O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=50)
    res, _ = U.run_eval_case(cfg, 40, False)
    assert res["zscore1_err"] <= EMB_ATOL, res["zscore1_err"]
    assert res["zscore_err"] <= EMB_ATOL


def test_pair_plane_saturation_is_flagged():
    """fp16 pair planes clamp at 65504 (weights are stored * 2^8: |w| >= 255.9): the conversion kernels raise a sticky
    device flag instead of silently saturating, and the Python layer turns it into an exception."""
    cfg = O.CVAEConfig(z_dim=10)
    eng = U.make_engine(cfg, 16)
    st = O.init_state(cfg, seed=42)
    eng.load_named(st)
    x1, x2, labels, eps = U.case_inputs(cfg, 16, False)
    dev = eng.device
    args = (x1.to(dev), x2.to(dev), labels.to(dev), None, eps.to(dev), 0.5)
    eng.train_fwd_bwd(*args)
    assert eng.device_flags(clear=True) == 0
    eng.raise_on_flags()  # nothing to raise
    big = {k: (v * 1.0e5 if k.endswith("layer4.1.conv2.weight") else v) for k, v in st.items()}
    eng.load_named(big)
    eng.train_fwd_bwd(*args)
    flags = eng.device_flags(clear=False)
    assert flags != 0
    with pytest.raises(OverflowError):
        eng.raise_on_flags()
    eng.load_named(st)
    eng.train_fwd_bwd(*args)
    assert eng.device_flags(clear=True) == 0  # sticky flag was cleared by raise_on_flags


def test_out_of_range_labels_raise_index_error():
    """nn.Embedding raises IndexError for a label outside its table (reference hippie/model.py:425-426).  CPU label
    tensors are checked on the host before the launch; device tensors are checked by the head kernels (flag, the row is
    skipped, neighbouring parameters are never read or written)."""
    from hippie_b200 import model as M
    torch.manual_seed(0)
    m = M.MultiModalCVAE(10, 50, 100, 5, 5, 3, max_batch=16).to("cuda:0")
    x1, x2 = torch.randn(8, 1, 50), torch.randn(8, 1, 100)
    src = torch.tensor([0, 1, 2, 3, 4, 0, 1, 5])  # 5 is outside [0, 5)
    with pytest.raises(IndexError):
        m(x1, x2, src)
    with pytest.raises(IndexError):
        m(x1, x2, torch.zeros(8, dtype=torch.int64), torch.tensor([0, 1, 2, 3, 0, 1, 2, 0]))  # class 3 outside [0, 3)
    with pytest.raises(IndexError):
        m(x1, x2, torch.tensor([0, -1, 0, 0, 0, 0, 0, 0]))
    # device-side labels: flagged by the kernel, nothing outside the tables is touched
    eng = m.engine
    before = eng.flat_params.clone()
    eng.flat_grads.zero_()
    eps = torch.zeros(8, 10, device="cuda:0")
    eng.train_fwd_bwd(x1.cuda(), x2.cuda(), src.cuda(), None, eps, 0.5)
    with pytest.raises(IndexError):
        eng.raise_on_flags()
    assert torch.equal(eng.flat_params, before)
    assert torch.isfinite(eng.flat_grads).all()


@pytest.mark.parametrize("B", [37, 512])
def test_no_kernel_writes_outside_its_buffers(B):
    """compute-sanitizer is closed on the pool this repository is built on, so the out-of-bounds check is our own:
    every buffer handed to hippie_bind (parameters, gradients, AdamW moments, BatchNorm buffers, the workspace) sits
    between canary bands; a full train step (ragged batch: B = 37 of max 64 leaves partial tiles everywhere), the
    optimizer, an eval forward and an embedding pass must leave every canary element untouched."""
    from hippie_b200.engine import Engine
    cfg = O.CVAEConfig(z_dim=10, num_classes=4)
    cap = 64 if B < 64 else B
    eng = Engine(cfg.z_dim, 50, 100, 5, cfg.num_sources, cfg.num_classes, True, cap).allocate("cuda:0", guard=1 << 16)
    eng.load_named(U.perturbed_state(cfg))
    x1, x2, labels, eps = U.case_inputs(cfg, B, True, seed=21)
    dev = eng.device
    cls, src = labels.unbind(1)
    args = (x1.to(dev), x2.to(dev), src.contiguous().to(dev), cls.contiguous().to(dev))
    for _ in range(3):  # eager, capture, replay
        s, _ = eng.train_fwd_bwd(*args, eps.to(dev), 0.5)
        eng.clip_adamw(1e-3, 0.01, 1, max_norm=1.0, step_cls=1, has_cls_grad=True)
    eng.embed(*args, zscore_ddof=1)
    eng.check_guards()
    assert torch.isfinite(s[:4]).all() and torch.isfinite(eng.flat_params).all()


def test_free_running_epoch_inside_reference_envelope(golden_dir):
    """BASELINE.json configs[1]/[2]: 200 free-running bs512 pretrain steps (no teacher forcing) against the curves of the
    reference's own classes {fp32 N threads, fp32 one thread, fp64} frozen in tests/golden/free_run_bs512.npz
    (tools/free_run_reference.py).  The reference is chaotic after step 0 (its own runs drift apart by up to ~1 % within
    ten steps and re-converge), so the engine has to stay within a small multiple of that spread, taken over a +-5 step
    window, and must end where the reference ends."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(golden_dir), "..", "tools"))
    import free_run_report as FR
    fx = np.load(os.path.join(golden_dir, "free_run_bs512.npz"))
    got = FR.engine_curve(fx)
    ref = fx["f64"][:, 0]
    spread, win = FR.envelope(fx)
    dev = np.abs(got[:, 0] - ref) / np.abs(ref)
    assert dev[0] <= LOSS_RTOL, dev[0]                         # identical state: the per-step bound
    assert (dev <= 6.0 * win + 3e-3).all(), (int(np.argmax(dev - 6 * win)), dev.max(), win.max())
    tail_ref, tail_got = ref[-20:].mean(), got[-20:, 0].mean()
    assert abs(tail_got - tail_ref) <= 5e-3 * tail_ref, (tail_got, tail_ref)
    assert got[-1, 0] < 0.2 * got[0, 0]                        # it trains


def test_weight_planes_kept_by_adamw_equal_a_fresh_conversion():
    """hippie_clip_adamw keeps the fp16 pair planes of the parameter buffer current (the per-step 64 MB conversion pass is
    gone).  After some steps (eager, then graph replays) a deterministic forward pass over the kept planes must equal,
    bit for bit, the same pass after hippie_params_changed has forced a conversion of the whole buffer."""
    cfg = O.CVAEConfig(z_dim=10, num_classes=4)
    st = U.perturbed_state(cfg)
    x1, x2, labels, eps = U.case_inputs(cfg, 48, True)
    eng = U.make_engine(cfg, 48)
    eng.load_named(st)
    dev = eng.device
    lab = labels.to(dev)
    io = (x1.to(dev), x2.to(dev), lab[:, 1].contiguous())
    for labelled in (False, True):  # label-free steps skip class_embedding in AdamW: its planes must stay valid too
        cls = lab[:, 0].contiguous() if labelled else None
        for i in range(6):
            s, _ = eng.train_fwd_bwd(io[0], io[1], io[2], cls, eps.to(dev), 0.5)
            eng.clip_adamw(1e-3, 0.01, i + 1, max_norm=1.0, scalars=s, step_cls=i + 1 if labelled else 0,
                           has_cls_grad=labelled)
        kept = eng.eval_forward(io[0], io[1], io[2], lab[:, 0].contiguous(), eps.to(dev))
        kept = {k: v.clone() for k, v in kept.items()}
        eng.params_changed()
        fresh = eng.eval_forward(io[0], io[1], io[2], lab[:, 0].contiguous(), eps.to(dev))
        for k in kept:
            assert torch.equal(kept[k], fresh[k]), (k, labelled)
    assert eng.device_flags(clear=True) == 0


def test_parameter_writes_behind_the_engine_are_seen():
    """In-place torch writes to the parameter buffer (load_state_dict, init functions, a broadcast) bump its version
    counter and the engine converts the planes again; writes through `.data` bypass the counter and need
    Engine.params_changed() -- the documented contract of hippie_params_changed."""
    cfg = O.CVAEConfig(z_dim=10)
    st = U.perturbed_state(cfg)
    x1, x2, labels, eps = U.case_inputs(cfg, 16, False)
    eng = U.make_engine(cfg, 16)
    eng.load_named(st)
    dev = eng.device
    io = (x1.to(dev), x2.to(dev), labels.to(dev), None)
    ref0 = eng.embed(*io)["mu"].clone()
    assert torch.equal(eng.embed(*io)["mu"], ref0)  # planes reused
    w = eng.named_state()["encoder_mod1.layer4.1.conv2.weight"]
    with torch.no_grad():
        w.mul_(1.5)  # version counter moves: detected
    ref1 = eng.embed(*io)["mu"].clone()
    assert not torch.equal(ref1, ref0)
    fresh = U.make_engine(cfg, 16)
    fresh.flat_params.copy_(eng.flat_params)
    fresh.bn_mean.copy_(eng.bn_mean), fresh.bn_var.copy_(eng.bn_var)
    assert torch.equal(fresh.embed(*io)["mu"], ref1)
    w.data.mul_(1.0 / 1.5)  # bypasses the counter: stale planes until reported
    if os.environ.get("HIPPIE_B200_KEEP_PLANES", "1") != "0":  # (the switch converts the planes in every call)
        assert torch.equal(eng.embed(*io)["mu"], ref1)
    eng.params_changed()
    fresh.flat_params.copy_(eng.flat_params)
    assert torch.equal(eng.embed(*io)["mu"], fresh.embed(*io)["mu"])


def test_shortcut_helper_stream_gives_the_serial_results(monkeypatch):
    """Batches <= 128 run the shortcut convolution of every down / up block (and its dgrad) on a helper stream beside the
    main path.  Against the same engine with the shortcuts kept on the chain (HIPPIE_B200_SHORTCUT_STREAM=0):
    training-mode and eval-mode forward outputs bit for bit (eager and graph replay), gradients to accumulation-order
    noise of the split-K weight gradients."""
    cfg = O.CVAEConfig(z_dim=10, num_classes=4)
    st = U.perturbed_state(cfg)
    x1, x2, labels, eps = U.case_inputs(cfg, 48, True)
    res = []
    for limit in ("0", "128"):
        monkeypatch.setenv("HIPPIE_B200_SHORTCUT_STREAM", limit)
        eng = U.make_engine(cfg, 48)
        eng.load_named(st)
        dev = eng.device
        lab = labels.to(dev)
        a = (x1.to(dev), x2.to(dev), lab[:, 1].contiguous(), lab[:, 0].contiguous(), eps.to(dev))
        outs = []
        for _ in range(3):  # eager, eager, replay
            ev = eng.eval_forward(*a)
            tr = eng.train_forward(*a)
            s, o = eng.train_fwd_bwd(*a, 0.5, outputs=True)
            outs.append(({k: v.clone() for k, v in ev.items()}, {k: v.clone() for k, v in tr.items()}, s.clone(),
                         {k: v.clone() for k, v in o.items()}, eng.flat_grads.clone()))
        res.append(outs)
        assert eng.device_flags(clear=True) == 0
    for serial, helper in zip(res[0], res[1]):
        for k in serial[0]:
            assert torch.equal(serial[0][k], helper[0][k]), ("eval", k)
            assert torch.equal(serial[1][k], helper[1][k]), ("train", k)
        assert torch.equal(serial[2][:4], helper[2][:4])
        for k in serial[3]:
            assert torch.equal(serial[3][k], helper[3][k]), ("step", k)
        assert U.rel_l2(helper[4], serial[4]) < 1e-5
