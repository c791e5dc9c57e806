"""Parity tests proper: the CUDA engine, called through the C ABI, against the CPU oracle on the same seeded
inputs (teacher-forced), against the golden fixtures frozen from the reference's own
classes, and size-independent properties at the benchmark size (bs512)."""
import os

import numpy as np
import pytest
import torch

from oracle import cvae_oracle as O
import parity_util as U

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5   # north_star: per-step loss within 1e-5 relative in fp32
EMB_ATOL = 1e-4    # north_star: embeddings within 1e-4 absolute

CASES = {
    "mm_z10_b48": (O.CVAEConfig(z_dim=10), 48, False),
    "mm_z10_b64_labelled": (O.CVAEConfig(z_dim=10, num_classes=4), 64, True),
    "mm_z32_b24": (O.CVAEConfig(z_dim=32), 24, False),
    "mm_z64_b16_labelled": (O.CVAEConfig(z_dim=64, num_classes=4), 16, True),
    "uni_wave_b24": (O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=50), 24, False),
    "uni_isi_b24_labelled": (O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=100, num_classes=4), 24, True),
    "mm_z10_b130_ragged_tiles": (O.CVAEConfig(z_dim=10), 130, False),
    "mm_z10_b2_minimum": (O.CVAEConfig(z_dim=10), 2, False),
    "mm_z10_b256_two_samples_per_head_cta": (O.CVAEConfig(z_dim=10), 256, False),
    "uni_wave_b300_labelled": (O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=50, num_classes=4), 300, True),
}


def _check_train(res, lr=1e-3, ratio=3.0):
    assert max(res["loss_rel"]) <= LOSS_RTOL, res["loss_rel"]
    for k, (e, r, scale) in res["out_err"].items():
        assert e <= max(10 * r, 2e-5 * max(scale, 1.0)), (k, e, r)
    for k, (e, r) in res["tap_err"].items():
        assert e <= max(10 * r, 2e-5), (k, e, r)
    # Gradients: per tensor err(engine, fp64) <= 2 x err(reference fp32, fp64) with the 1e-6 * ||g|| floor,
    # flat relative L2 <= 2e-3 -- with the LeakyReLU branches TEACHER-FORCED: a backbone activation whose input is ~0 can
    # land on the other side of zero in any fp32 evaluation order (the reference's own fp32 run does, "flips" columns of
    # profiles/r02_grad_error_table.md); that one element then scales its gradient path by 100 (slope 0.01 vs 1) and
    # moves whole tensors by 1e-3..1e-2 whatever the arithmetic.  So the fp64 oracle is re-run with every such branch set
    # to the one the compared implementation took (oracle `masks`), which leaves pure arithmetic error: measured worst
    # ratio 0.9 .. 2.0 (tcgen05 pair planes) and 1.3 .. 2.8 (FP32 CUDA cores) over all cases, flat 2e-6 .. 6e-6.  The
    # bound is 3 x (the ratio of two rounding-noise norms scatters), not the 10 x of round 1.
    gn = res["grad_global_norm"]
    for n, (e, r, nn) in res["grad_err_tf"].items():
        assert e <= ratio * r + 1e-6 * gn, (n, e, r, nn)
    assert res["grad_flat_rel_tf"] <= min(2e-3, ratio * res["grad_flat_rel_f32_tf"] + 1e-6), \
        (res["grad_flat_rel_tf"], res["grad_flat_rel_f32_tf"])
    assert res["loss_rel_tf"] <= LOSS_RTOL
    # ... and free-running (branches as each side took them): a gross-error bound (a wrong layout or a missing term
    # would be O(1)); tight when neither side flipped
    clean = res["flips_eng"] == 0 and res["flips_f32"] == 0
    for n, (e, r, nn) in res["grad_err"].items():
        bound = (ratio * r + 1e-6 * gn) if clean else (10 * r + 0.05 * nn)
        assert e <= bound + 1e-6 * gn, (n, e, r, nn, res["flips_eng"], res["flips_f32"])
    assert res["grad_flat_rel"] <= (ratio * res["grad_flat_rel_f32"] + 1e-6 if clean else 0.03), \
        (res["grad_flat_rel"], res["grad_flat_rel_f32"], res["flips_eng"], res["flips_f32"])
    assert res["no_grad_params"] == []
    assert res["running_err"] <= 1e-5
    assert abs(res["grad_norm_eng"] - res["grad_norm_f64"]) <= 5e-3 * res["grad_norm_f64"]
    assert abs(res["clip_eng"] - res["clip_f64"]) <= 5e-3 * res["clip_f64"]
    assert res["param_abs_err"] <= 2 * lr + 1e-6  # Adam's first step is ~lr*sign(g): noise-level grads may flip
    assert res["exp_avg_rel"] <= 0.03
    if "cls_emb_untouched" in res:
        assert res["cls_emb_untouched"]  # torch skips params whose grad is None (label-free steps)


@pytest.mark.parametrize("name", list(CASES))
def test_train_step_matches_oracle(name):
    cfg, B, lab = CASES[name]
    res, eng = U.run_train_case(cfg, B, lab)
    # B = 2: the head's BatchNorm1d normalises two samples (x_hat = +-1, invstd = 2 / |x1 - x2|), which amplifies rounding
    # by whatever |x1 - x2| happens to be; the ratio of two such noise norms scatters more (measured 4.7)
    _check_train(res, ratio=10.0 if B == 2 else 3.0)


def test_train_step_large_dynamic_range():
    """Inputs with the dynamic range of the real cellexplorer rows (index column leak: values up to ~400)."""
    res, _ = U.run_train_case(O.CVAEConfig(z_dim=10), 48, False, real_scale=True)
    assert max(res["loss_rel"]) <= LOSS_RTOL, res["loss_rel"]


@pytest.mark.parametrize("name", ["mm_z10_b48", "mm_z10_b64_labelled", "mm_z32_b24", "uni_wave_b24", "uni_isi_b24_labelled"])
def test_eval_forward_and_embedding_match_oracle(name):
    cfg, B, lab = CASES[name]
    res, _ = U.run_eval_case(cfg, B, lab)
    for k, e in res["abs_err"].items():
        assert e <= EMB_ATOL, (k, e)
    for k, e in res["emb_err"].items():
        assert e <= EMB_ATOL, (k, e)
    assert res["zscore_err"] <= 1e-4
    assert max(res["loss_rel"]) <= LOSS_RTOL


GOLDEN = {
    "real48_z10": O.CVAEConfig(z_dim=10),
    "synth64_labelled_z10": O.CVAEConfig(z_dim=10, num_classes=4),
    "synth24_z32": O.CVAEConfig(z_dim=32),
    "uni_wave24_z10": O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=50),
    "uni_isi24_z10": O.CVAEConfig(z_dim=10, multimodal=False, output_size_wave=100),
}


@pytest.mark.parametrize("tag", list(GOLDEN))
def test_against_reference_golden_fixtures(golden_dir, tag):
    """Fixtures were produced by the reference's own classes (oracle/make_golden.py): seed-42 init, first
    train-mode forward and first optimisation step."""
    from hippie_b200 import model as M
    cfg, fx = GOLDEN[tag], np.load(os.path.join(golden_dir, tag + ".npz"))
    lr, wd, beta, w1, w2, clip = [float(v) for v in fx["hyper"]]
    dev = torch.device("cuda:0")
    torch.manual_seed(42)  # the module mirror must reproduce the reference's initial parameters bit for bit
    if cfg.multimodal:
        m = M.MultiModalCVAE(cfg.z_dim, 50, 100, 5, cfg.num_sources, cfg.num_classes, max_batch=64)
        tm = M.MultiModalCVAETrainModule(m, learning_rate=lr, weight_decay=wd, beta=beta, mod1_weight=w1, mod2_weight=w2)
    else:
        m = M.hippieUnimodalCVAE(cfg.z_dim, cfg.output_size_wave, 5, cfg.num_sources, cfg.num_classes, max_batch=64)
        tm = M.hippieUnimodalEmbeddingModelCVAE(m, learning_rate=lr, weight_decay=wd, beta=beta)
    tm.to(dev)
    x1 = torch.tensor(fx["x1"])
    x2 = torch.tensor(fx["x2"]) if "x2" in fx else None
    labels = torch.tensor(fx["labels"])
    torch.manual_seed(int(fx["eps_seeds"][0]))
    eps = torch.randn(x1.shape[0], cfg.z_dim).to(dev)
    batch = (x1, x2, labels) if cfg.multimodal else (x1, labels)
    m.train()
    cls, src = (labels.unbind(1) if labels.dim() == 2 else (None, labels))
    outs = m(x1, x2, src, cls, eps=eps) if cfg.multimodal else m(x1, src, cls, eps=eps)
    keys = ["enc", "mu", "logvar", "dec1"] + (["dec2"] if cfg.multimodal else [])
    for k, o in zip(keys, outs):
        r = torch.tensor(fx[f"f32_fwd0_{k}"])
        tol = 5e-5 * max(1.0, r.abs().max().item())  # fp32 vs fp32: each side sits ~2e-5 from the fp64 value
        assert (o.cpu().reshape(r.shape) - r).abs().max().item() <= tol, k
    # undo the running-statistics update of the probe forward, then take the reference's first step
    ref0 = O.init_state(cfg, seed=42)
    m.load_state_dict(ref0)
    loss = tm.training_step(batch, 0, eps=eps)
    tm.optimizer.step(max_norm=clip)
    got = tm._ring[tm._ring_pos - 1].cpu().numpy()
    ref_loss = fx["f32_s0_loss"]
    sel = [0, 1, 2, 3] if cfg.multimodal else [0, 1, 3]
    np.testing.assert_allclose(got[sel], ref_loss[sel], rtol=LOSS_RTOL, atol=1e-9)
    assert float(loss) == pytest.approx(float(ref_loss[0]), rel=LOSS_RTOL)
    np.testing.assert_allclose(tm.optimizer.last_scalars[4].item(), float(fx["f32_s0_grad_norm"]), rtol=3e-3)
    # whole tensors through the fixture's checksums: per-tensor L2 norms of the reference's raw gradients
    gnames = [str(n) for n in fx["s0_grad_names"]]
    grads = m.engine.named_grads()
    ref_l2 = fx["f32_s0_grad_l2"]
    gtot = float(np.sqrt((ref_l2 ** 2).sum()))
    for n, r in zip(gnames, ref_l2):
        got_n = grads[n].double().norm().item()
        assert abs(got_n - r) <= 5e-3 * r + 1e-5 * gtot, (n, got_n, r)
    names = [str(n) for n in fx["param_names"]]
    sd = m.state_dict()
    # ... and the sums of the updated parameters (Adam's first step moves every element by ~lr: an element whose
    # noise-level gradient has the other sign is off by 2 lr, so the bound scales with the tensor size)
    for n, r in zip(names, fx["f32_s0_param_sum"]):
        t = sd[n].detach().double()
        assert abs(t.sum().item() - r) <= lr * (2.0 + 0.05 * t.numel()), (n, t.sum().item(), r)
    run_names = [str(n) for n in fx["running_names"]]
    run_head = np.stack([np.pad(sd[k].detach().cpu().double().flatten()[:6].numpy(), (0, max(0, 6 - sd[k].numel())))
                         for k in run_names])
    assert np.abs(run_head - fx["f32_s0_running_head"]).max() <= 1e-5 * max(1.0, np.abs(fx["f32_s0_running_head"]).max())
    head = np.stack([np.pad(sd[k].detach().cpu().double().flatten()[:6].numpy(), (0, max(0, 6 - sd[k].numel())))
                     for k in names])
    assert np.abs(head - fx["f32_s0_param_head"]).max() <= 2 * lr + 1e-7


def test_ragged_last_batch_does_not_see_stale_rows():
    """An engine sized for 64 runs 64 and then 37 samples; the second result must equal a fresh run of 37."""
    cfg = O.CVAEConfig(z_dim=10)
    eng = U.make_engine(cfg, 64)
    res64, _ = U.run_train_case(cfg, 64, False, engine=eng, seed=5)
    res37, _ = U.run_train_case(cfg, 37, False, engine=eng, seed=6)
    _check_train(res37)


def test_error_behaviour():
    cfg = O.CVAEConfig(z_dim=10)
    eng = U.make_engine(cfg, 8)
    dev = eng.device
    x1, x2 = torch.zeros(16, 1, 50, device=dev), torch.zeros(16, 1, 100, device=dev)
    src, eps = torch.zeros(16, dtype=torch.int64, device=dev), torch.zeros(16, 10, device=dev)
    with pytest.raises(RuntimeError, match="max_batch"):
        eng.train_fwd_bwd(x1, x2, src, None, eps, 0.5)
    with pytest.raises(RuntimeError, match="eps"):
        eng.train_fwd_bwd(x1[:8], x2[:8], src[:8], None, None, 0.5)
    inf = U.make_engine(cfg, 8, inference_only=True)
    with pytest.raises(RuntimeError, match="inference-only"):
        inf.train_fwd_bwd(x1[:8], x2[:8], src[:8], None, eps[:8], 0.5)


def test_properties_at_benchmark_size_bs512():
    """Size-independent properties at BASELINE.json's batch size: loss identity, determinism of the forward,
    batch independence of the eval-mode embedding, zero padding rows, clip invariant."""
    cfg = O.CVAEConfig(z_dim=10)
    B = 512
    eng = U.make_engine(cfg, B)
    st = O.init_state(cfg, seed=42)
    eng.load_named(st)
    dev = eng.device
    x1, x2, labels, eps = U.case_inputs(cfg, B, False, seed=77)
    x1, x2, src, eps = x1.to(dev), x2.to(dev), labels.to(dev), eps.to(dev)
    beta, w1, w2 = 0.5, 0.7, 1.3
    s1, _ = eng.train_fwd_bwd(x1, x2, src, None, eps, beta, w1, w2)
    g1 = eng.flat_grads.clone()
    a = s1.cpu().double()
    assert abs(a[0] - (w1 * a[1] + w2 * a[2] + beta * a[3])) <= 1e-6 * abs(a[0])
    eng.load_named(st)  # restore running statistics
    s2, _ = eng.train_fwd_bwd(x1, x2, src, None, eps, beta, w1, w2)
    assert torch.equal(s1[:4], s2[:4])  # forward is deterministic (no atomics on the forward path)
    rel = ((eng.flat_grads - g1).norm() / g1.norm()).item()
    assert rel <= 1e-5  # wgrad split-K uses fp32 atomics: order-dependent rounding only
    # every padding row of every activation / gradient tensor is still zero
    ws = eng.workspace.view(torch.float32)
    for t in eng.tensors:
        rows = t.L + 2
        v = ws[t.offset:t.offset + B * rows * t.C].view(B, rows, t.C)
        assert v[:, 0].abs().max().item() == 0.0 and v[:, -1].abs().max().item() == 0.0, t.name
    # clip invariant: after clipping the applied gradient has norm <= max_norm
    sc = eng.clip_adamw(1e-3, 0.01, 1, max_norm=1.0)
    norm, coef = sc[4].item(), sc[5].item()
    assert abs(norm - g1.norm().item()) <= 1e-4 * norm  # the device's norm is the norm of the gradient buffer
    assert coef == pytest.approx(min(1.0, 1.0 / (norm + 1e-6)), rel=1e-5)
    # eval-mode embedding: units are independent -> chunked == whole, bit for bit (idempotence under re-batching)
    whole = eng.embed(x1, x2, src)
    halves = [eng.embed(x1[i:i + 256].contiguous(), x2[i:i + 256].contiguous(), src[i:i + 256].contiguous()) for i in (0, 256)]
    for k in ("enc", "mu", "logvar"):
        assert torch.equal(whole[k], torch.cat([h[k] for h in halves])), k
    z = eng.embed(x1, x2, src, zscore_ddof=0)["enc"]
    assert z.mean(dim=1).abs().max().item() <= 1e-5 and (z.var(dim=1, unbiased=False) - 1).abs().max().item() <= 1e-4


def test_split_step_equals_whole_step():
    """hippie_train_fwd_bwd_part(0) + (1) == hippie_train_fwd_bwd: same loss scalars bit for bit, same gradients up to the
    order of the weight-gradient atomics, and the documented halves of the gradient buffer are final after each part."""
    cfg = O.CVAEConfig(z_dim=10)
    B = 96
    eng = U.make_engine(cfg, B)
    st = O.init_state(cfg, seed=42)
    dev = eng.device
    x1, x2, labels, eps = U.case_inputs(cfg, B, False, seed=5)
    x1, x2, src, eps = x1.to(dev), x2.to(dev), labels.to(dev), eps.to(dev)
    for _ in range(3):  # eager call, graph capture, graph replay
        eng.load_named(st)
        s_whole, _ = eng.train_fwd_bwd(x1, x2, src, None, eps, 0.5, 1.0, 1.0)
        g_whole = eng.flat_grads.clone()
        eng.load_named(st)
        s0 = eng.train_fwd_bwd_part(0, x1, x2, src, None, eps, 0.5, 1.0, 1.0)
        split = eng.grad_split
        tail = eng.flat_grads[split:].clone()
        assert eng.flat_grads[:split].abs().max().item() == 0.0  # encoder gradients untouched so far
        eng.train_fwd_bwd_part(1, x1, x2, src, None, eps, 0.5, 1.0, 1.0)
        assert torch.equal(s_whole[:4], s0[:4])
        assert torch.equal(eng.flat_grads[split:], tail)  # part 1 does not touch the decoder / head gradients
        rel = ((eng.flat_grads - g_whole).norm() / g_whole.norm()).item()
        assert rel <= 1e-5, rel
    assert 0 < split < eng.param_floats
    assert eng.params[[p.offset for p in eng.params].index(split)].name == "fusion_encoder.0.weight"
    # parts 0, 2, 3: the encoder backward in its deep and shallow halves (hippie_grad_bounds)
    bounds = eng.grad_bounds
    names = {p.offset: p.name for p in eng.params}
    assert [names[b] for b, _, _ in bounds] == ["encoder_mod1.conv1.weight", "encoder_mod2.conv1.weight"]
    assert [names[d] for _, d, _ in bounds] == ["encoder_mod1.layer3.0.conv1.weight", "encoder_mod2.layer3.0.conv1.weight"]
    assert bounds[0][2] == bounds[1][0] and bounds[1][2] == split
    for _ in range(3):
        eng.load_named(st)
        s0 = eng.train_fwd_bwd_part(0, x1, x2, src, None, eps, 0.5, 1.0, 1.0)
        eng.train_fwd_bwd_part(2, x1, x2, src, None, eps, 0.5, 1.0, 1.0)
        deep = [eng.flat_grads[d:e].clone() for _, d, e in bounds]
        for b, d, e in bounds:
            assert eng.flat_grads[b:d].abs().max().item() == 0.0  # shallow halves untouched so far
            assert eng.flat_grads[d:e].abs().max().item() > 0.0
        eng.train_fwd_bwd_part(3, x1, x2, src, None, eps, 0.5, 1.0, 1.0)
        for (b, d, e), saved in zip(bounds, deep):
            assert torch.equal(eng.flat_grads[d:e], saved)  # part 3 does not touch the deep halves
        assert torch.equal(s_whole[:4], s0[:4])
        rel = ((eng.flat_grads - g_whole).norm() / g_whole.norm()).item()
        assert rel <= 1e-5, rel


def test_fp32_cuda_core_path_is_the_yardstick():
    """conv_path = 1 (FP32 CUDA-core implicit GEMMs, no pair planes) passes the same parity gate, and the tensor-core
    path agrees with it on the forward to fp32 rounding level."""
    cfg, B, lab = CASES["mm_z10_b48"]
    res1, eng1 = U.run_train_case(cfg, B, lab, conv_path=1)
    assert eng1.conv_path_in_use() == 1
    _check_train(res1)
    res0, eng0 = U.run_train_case(cfg, B, lab, conv_path=0)
    assert eng0.conv_path_in_use() == 2
    for a, b in zip(res0["loss_rel"], res1["loss_rel"]):
        assert abs(a - b) <= 2e-6


def test_free_running_steps_stay_inside_the_reference_envelope():
    """Six consecutive optimisation steps WITHOUT teacher forcing.  The reference itself is chaotic after step 0: its own
    fp32 and fp64 runs drift apart by 1e-4 .. 1e-2 relative within a few steps (Adam's first updates
    are ~lr * sign(g), so rounding-level gradient differences flip updates).  The engine's loss curve therefore has to
    stay within a small multiple of the reference's OWN fp32-vs-fp64 spread, measured in the same test."""
    cfg = O.CVAEConfig(z_dim=10)
    B, steps = 256, 6
    eng = U.make_engine(cfg, B)
    st32 = O.init_state(cfg, seed=42)
    st64 = U.to_dtype(st32, torch.float64)
    eng.load_named(st32)
    dev = eng.device
    x1, x2, labels, g = O.synthetic_batch(B * steps, seed=321)
    eps = torch.randn(steps, B, cfg.z_dim, generator=g)
    opt32, opt64 = O.new_opt_state(st32, cfg), O.new_opt_state(st64, cfg)
    ref32, ref64, got = [], [], []
    kw = dict(lr=1e-3, weight_decay=0.01, beta=0.5, max_norm=1.0)
    for i in range(steps):
        sl = slice(i * B, (i + 1) * B)
        st32, opt32, info = O.train_step(st32, opt32, cfg, x1[sl], x2[sl], labels[sl], eps[i], **kw)
        ref32.append(float(info["loss"]))
        st64, opt64, info = O.train_step(st64, opt64, cfg, x1[sl].double(), x2[sl].double(), labels[sl], eps[i].double(), **kw)
        ref64.append(float(info["loss"]))
        s, _ = eng.train_fwd_bwd(x1[sl].to(dev), x2[sl].to(dev), labels[sl].to(dev), None, eps[i].to(dev), 0.5)
        eng.clip_adamw(1e-3, 0.01, i + 1, max_norm=1.0)
        got.append(float(s[0]))
    assert abs(got[0] - ref64[0]) <= 1e-5 * abs(ref64[0])  # step 0 starts from identical state
    spread = 0.0
    for a, r32, r64 in zip(got, ref32, ref64):
        spread = max(spread, abs(r32 - r64) / abs(r64))
        # |fp32 - fp64| of the reference is ONE draw of the divergence; the engine's curve is another one (different
        # summation orders, atomics in the weight gradients), so it gets a multiple of it.  Teacher-forced parity is
        # checked tightly above; a real error shows up as O(1e-1) here.
        assert abs(a - r64) / abs(r64) <= 8.0 * spread + 2e-3, (got, ref32, ref64)
    assert got[-1] < got[0]


def test_device_preprocessing_matches_the_host_dataset(golden_dir):
    """hippie_preprocess_batch (GPU) vs EphysDataset.__getitem__ semantics (oracle.dataset_item, pinned against the
    reference's fixture): waveforms bit for bit; ISI within 2 ulp (ATen's CPU logf is not correctly rounded)."""
    from hippie_b200.dataloading import DeviceTable
    fx = np.load(os.path.join(golden_dir, "cellexplorer_raw48.npz"))
    tab = DeviceTable(fx["wf"], fx["isi"], labels=np.arange(48) % 4)
    idx = [5, 0, 47, 13, 13, 2]
    x1, x2, lab = tab.batch(idx)
    assert x1.shape == (6, 1, 50) and x2.shape == (6, 1, 100) and lab.tolist() == [i % 4 for i in idx]
    assert np.array_equal(x1.cpu().numpy(), fx["x1"][idx])  # the reference's own output, bit for bit
    ref2 = torch.tensor(fx["x2"][idx])
    ulp = (x2.cpu().view(torch.int32) - ref2.view(torch.int32)).abs().max().item()
    assert ulp <= 2, ulp
    assert (x2.cpu() != ref2).float().mean().item() < 0.02
    # ragged widths (other source tables are 40 / 351 samples wide) and the identity index
    rng = np.random.default_rng(1)
    for ww, wi in ((40, 51), (351, 100), (50, 100)):
        wf, isi = rng.normal(size=(9, ww)), np.abs(rng.normal(size=(9, wi)))
        a, b = DeviceTable(wf, isi).batch(range(9))
        for r in range(9):
            oa, ob = O.dataset_item(wf[r], isi[r])
            assert torch.equal(a[r].cpu(), oa), (ww, r)
            assert (b[r].cpu().view(torch.int32) - ob.view(torch.int32)).abs().max().item() <= 2
    e1, e2 = DeviceTable(np.zeros((3, 47)), np.zeros((3, 100))).batch([])
    assert e1.shape == (0, 1, 50) and e2.shape == (0, 1, 100)


def test_large_batch_embedding_equals_small_batches():
    """4096 units in one call (multi-wave grids, the throughput configuration of bench.py's embedding workload) must
    equal, bit for bit, the concatenation of bs512 calls: units are independent in eval mode."""
    cfg = O.CVAEConfig(z_dim=10)
    N = 4096
    eng = U.make_engine(cfg, N, inference_only=True)
    eng.load_named(U.perturbed_state(cfg))
    x1, x2, labels, _ = O.synthetic_batch(N, seed=9)
    dev = eng.device
    x1, x2, src = x1.to(dev), x2.to(dev), labels.to(dev)
    whole = eng.embed(x1, x2, src)
    parts = [eng.embed(x1[i:i + 512].contiguous(), x2[i:i + 512].contiguous(), src[i:i + 512].contiguous())
             for i in range(0, N, 512)]
    for k in ("enc", "mu", "logvar"):
        assert torch.equal(whole[k], torch.cat([p[k] for p in parts])), k
