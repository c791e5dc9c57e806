"""CPU oracle for the HIPPIE cVAE hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain, functional restatement (torch CPU tensors, no nn.Module, no
Lightning) of the arithmetic the reference performs on the path named by
BASELINE.json:north_star.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and only as the
checker / CPU baseline -- never as the product path (the product path is
`hippie_b200/` + `libhippie_b200.so` and fails loudly without CUDA).

Parity pinning: the reference holds NO golden vectors for this path.
The oracle is therefore pinned against outputs of the reference's own classes executed
in the build container: `oracle/make_golden.py` imports `/root/reference/hippie` (with
the `pytorch_lightning` stub in `oracle/_plstub`), checks this restatement against it and
freezes the results into `tests/golden/*.npz`; `tests/test_oracle_golden.py` re-checks the
oracle against those fixtures everywhere (no `/root/reference` needed at test time).

Reference anchors (all paths relative to /root/reference):
  * ResizeConv1d / BasicBlockEnc / BasicBlockDec / ResNet18Enc / ResNet18Dec
      hippie/backbones.py:6-16, 19-41, 44-70, 73-103, 106-141
  * MultiModalCVAE                      hippie/model.py:350-432
  * hippieUnimodalCVAE                  hippie/model.py:12-72
  * loss (MSE x2 + beta * KL)           hippie/model.py:454-482 (multimodal), 93-115 (unimodal)
  * AdamW construction                  hippie/model.py:447 (torch.optim.AdamW defaults)
  * gradient clipping                   scripts/train_model_with_multimodal.py:55,701
                                        (Lightning gradient_clip_val -> clip_grad_norm_)
  * dataset transform                   hippie/dataloading.py:27-56
  * embedding extraction                scripts/train_model_with_multimodal.py:22-34,
                                        scripts/utils.py:75-101
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm1d default
BN_MOMENTUM = 0.1      # nn.BatchNorm1d default
SLOPE_BACKBONE = 0.01  # F.leaky_relu default (hippie/backbones.py:37,40,66,69,95)
SLOPE_HEAD = 0.2       # nn.LeakyReLU(0.2)  (hippie/model.py:367,381,384,388,391)


@dataclass(frozen=True)
class CVAEConfig:
    """Constructor arguments of MultiModalCVAE (hippie/model.py:352) or, with
    multimodal=False, hippieUnimodalCVAE (hippie/model.py:13; output_size_wave is then
    the single `output_size`)."""
    z_dim: int = 10
    output_size_wave: int = 50
    output_size_isi: int = 100
    class_hidden_dim: int = 5
    num_sources: int = 5
    num_classes: int = 5
    multimodal: bool = True


# --------------------------------------------------------------------------------------
# Parameter / buffer inventory in construction order (== state_dict order)
# --------------------------------------------------------------------------------------
def _conv(spec, name, cout, cin, k, bias):
    spec.append((name + ".weight", "conv_w", (cout, cin, k)))
    if bias:
        spec.append((name + ".bias", "conv_b", (cout, cin, k)))


def _bn(spec, name, c):
    spec.append((name + ".weight", "ones", (c,)))
    spec.append((name + ".bias", "zeros", (c,)))
    spec.append((name + ".running_mean", "buf_zeros", (c,)))
    spec.append((name + ".running_var", "buf_ones", (c,)))
    spec.append((name + ".num_batches_tracked", "buf_count", ()))


def _linear(spec, name, nout, nin):
    spec.append((name + ".weight", "lin_w", (nout, nin)))
    spec.append((name + ".bias", "lin_b", (nout, nin)))


def _enc_spec(spec, p, z):
    # ResNet18Enc.__init__ (backbones.py:74-84); BasicBlockEnc.__init__ (:20-34)
    _conv(spec, p + ".conv1", 64, 1, 3, False)
    _bn(spec, p + ".bn1", 64)
    in_planes = 64
    for li, (planes, stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
        for bi, s in enumerate([stride, 1]):
            q = f"{p}.layer{li}.{bi}"
            out = in_planes * s
            _conv(spec, q + ".conv1", out, in_planes, 3, False)
            _bn(spec, q + ".bn1", out)
            _conv(spec, q + ".conv2", out, out, 3, False)
            _bn(spec, q + ".bn2", out)
            if s != 1:
                _conv(spec, q + ".shortcut.0", out, in_planes, 1, False)
                _bn(spec, q + ".shortcut.1", out)
            in_planes = planes
    _linear(spec, p + ".linear", 2 * z, 512)


def _dec_spec(spec, p, z, output_size):
    # ResNet18Dec.__init__ (backbones.py:107-118); BasicBlockDec.__init__ (:45-63).
    # _make_layer reverses the strides, so the up-sampling block is LAST in each layer.
    _linear(spec, p + ".linear", 512, 2 * z)
    in_planes = 512
    for li, (planes, stride) in [(4, (256, 2)), (3, (128, 2)), (2, (64, 2)), (1, (64, 1))]:
        for bi, s in enumerate([1, stride]):
            q = f"{p}.layer{li}.{bi}"
            out = in_planes // s
            _conv(spec, q + ".conv2", in_planes, in_planes, 3, False)
            _bn(spec, q + ".bn2", in_planes)
            if s == 1:
                _conv(spec, q + ".conv1", out, in_planes, 3, False)
                _bn(spec, q + ".bn1", out)
            else:
                _conv(spec, q + ".conv1.conv", out, in_planes, 3, True)
                _bn(spec, q + ".bn1", out)
                _conv(spec, q + ".shortcut.0.conv", out, in_planes, 3, True)
                _bn(spec, q + ".shortcut.1", out)
        in_planes = planes
    _conv(spec, p + ".conv1.conv", 1, 64, 3, True)
    _linear(spec, p + ".linear_out", output_size, 64)


def model_spec(cfg: CVAEConfig) -> List[Tuple[str, str, tuple]]:
    """[(state_dict key, init kind, shape-or-fan-shape)] in the reference's construction
    order, which is also its state_dict() order."""
    z, h = cfg.z_dim, cfg.class_hidden_dim
    spec: List[Tuple[str, str, tuple]] = []
    if cfg.multimodal:
        # MultiModalCVAE.__init__ (model.py:352-395)
        _enc_spec(spec, "encoder_mod1", z)
        _enc_spec(spec, "encoder_mod2", z)
        _linear(spec, "fusion_encoder.0", 2 * z, 4 * z + 2 * h)
        _bn(spec, "fusion_encoder.1", 2 * z)
        _linear(spec, "fusion_encoder.3", z, 2 * z)
        spec.append(("source_embedding.weight", "normal", (cfg.num_sources, h)))
        spec.append(("class_embedding.weight", "normal", (cfg.num_classes, h)))
        _linear(spec, "z_mean", z, z)
        _linear(spec, "z_log_var", z, z)
        for m in ("decoder_fc_mod1", "decoder_fc_mod2"):
            _linear(spec, m + ".0", 2 * z, z + 2 * h)
            _linear(spec, m + ".2", 2 * z, 2 * z)
            _bn(spec, m + ".3", 2 * z)
        _dec_spec(spec, "decoder_mod1", z, cfg.output_size_wave)
        _dec_spec(spec, "decoder_mod2", z, cfg.output_size_isi)
    else:
        # hippieUnimodalCVAE.__init__ (model.py:13-44)
        _enc_spec(spec, "encoder", z)
        _linear(spec, "encoder_fc.0", 2 * z, 2 * z + 2 * h)
        _bn(spec, "encoder_fc.1", 2 * z)
        _linear(spec, "encoder_fc.3", z, 2 * z)
        _bn(spec, "encoder_fc.4", z)
        spec.append(("source_embedding.weight", "normal", (cfg.num_sources, h)))
        spec.append(("class_embedding.weight", "normal", (cfg.num_classes, h)))
        _linear(spec, "z_mean", z, z)
        _linear(spec, "z_log_var", z, z)
        _linear(spec, "decoder_fc.0", 2 * z, z + 2 * h)
        _linear(spec, "decoder_fc.2", 2 * z, 2 * z)
        _bn(spec, "decoder_fc.3", 2 * z)
        _dec_spec(spec, "decoder", z, cfg.output_size_wave)
    return spec


def _torch_shape(kind, shp):
    if kind == "conv_b" or kind == "lin_b":
        return (shp[0],)
    return tuple(shp)


def is_buffer(kind: str) -> bool:
    return kind.startswith("buf_")


def param_names(cfg: CVAEConfig) -> List[str]:
    return [n for n, k, _ in model_spec(cfg) if not is_buffer(k)]


def init_state(cfg: CVAEConfig, seed: Optional[int] = None, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Restates torch's default initialisers in the reference's construction order so that,
    after the same `torch.manual_seed`, the values equal the reference model's bit for bit.
    nn.Conv1d / nn.Linear: kaiming_uniform_(a=sqrt(5)) on the weight, U(-1/sqrt(fan_in), ..)
    on the bias; nn.Embedding: N(0,1); BatchNorm1d: ones/zeros (no RNG)."""
    if seed is not None:
        torch.manual_seed(seed)
    st: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, kind, shp in model_spec(cfg):
        if kind in ("conv_w", "lin_w"):
            fan_in = shp[1] * (shp[2] if len(shp) == 3 else 1)
            gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
            std = gain / math.sqrt(fan_in)
            bound = math.sqrt(3.0) * std
            t = torch.empty(shp).uniform_(-bound, bound)
        elif kind in ("conv_b", "lin_b"):
            fan_in = shp[1] * (shp[2] if len(shp) == 3 else 1)
            bound = 1 / math.sqrt(fan_in)
            t = torch.empty(shp[0]).uniform_(-bound, bound)
        elif kind == "normal":
            t = torch.empty(shp).normal_()
        elif kind in ("ones", "buf_ones"):
            t = torch.ones(shp)
        elif kind in ("zeros", "buf_zeros"):
            t = torch.zeros(shp)
        elif kind == "buf_count":
            t = torch.tensor(0, dtype=torch.long)
        else:  # pragma: no cover
            raise ValueError(kind)
        st[name] = t if t.dtype == torch.long else t.to(dtype)
    return st


# --------------------------------------------------------------------------------------
# Layers (explicit formulas; contraction itself via F.conv1d / F.linear)
# --------------------------------------------------------------------------------------
class _Ctx:
    """Carries mode + collects BatchNorm running-stat updates (functional: the caller's
    state is never mutated; `new_buffers` holds what the reference would have written)."""

    def __init__(self, st, train: bool, masks=None):
        self.st = st
        self.train = train
        # test instrument (never set by the reference path): {activation name: bool [B,C,L]} forces the LeakyReLU branch
        # of that site, so that a comparison with an implementation whose rounding put a ~0 input on the other side of
        # zero measures arithmetic error instead of the factor-100 slope change
        self.masks = masks
        self.new_buffers: Dict[str, torch.Tensor] = {}
        self.taps: Dict[str, torch.Tensor] = {}  # named intermediates for per-layer parity tests


def _lrelu(x, slope, cx=None, site=None):
    pos = x > 0
    if cx is not None and cx.masks is not None and site in cx.masks:
        pos = cx.masks[site]
    return torch.where(pos, x, x * slope)


def _bn_apply(cx: _Ctx, name: str, x: torch.Tensor) -> torch.Tensor:
    """BatchNorm1d over [B,C,L] or [B,C].  Training: batch mean and BIASED variance
    normalise; running_var takes the UNBIASED variance; momentum 0.1; eps 1e-5."""
    st = cx.st
    dims = (0, 2) if x.dim() == 3 else (0,)
    shape = (1, -1, 1) if x.dim() == 3 else (1, -1)
    if cx.train:
        n = x.numel() // x.shape[1]
        mean = x.mean(dim=dims)
        var_b = ((x - mean.view(shape)) ** 2).mean(dim=dims)
        with torch.no_grad():
            var_u = var_b * (n / max(n - 1, 1))
            cx.new_buffers[name + ".running_mean"] = (
                (1 - BN_MOMENTUM) * st[name + ".running_mean"] + BN_MOMENTUM * mean.detach())
            cx.new_buffers[name + ".running_var"] = (
                (1 - BN_MOMENTUM) * st[name + ".running_var"] + BN_MOMENTUM * var_u.detach())
            cx.new_buffers[name + ".num_batches_tracked"] = st[name + ".num_batches_tracked"] + 1
    else:
        mean = st[name + ".running_mean"]
        var_b = st[name + ".running_var"]
    xhat = (x - mean.view(shape)) / torch.sqrt(var_b.view(shape) + BN_EPS)
    return xhat * st[name + ".weight"].view(shape) + st[name + ".bias"].view(shape)


def _up_nearest(x, s):
    """F.interpolate(x, scale_factor=s) default mode 'nearest': out[i] = in[i // s]."""
    L = x.shape[-1]
    idx = torch.arange(L * s) // s
    return x[..., idx]


def _conv1d(cx, name, x, stride, padding):
    return F.conv1d(x, cx.st[name + ".weight"], cx.st.get(name + ".bias"), stride=stride, padding=padding)


def _block_enc(cx: _Ctx, p: str, x, stride):
    # BasicBlockEnc.forward (backbones.py:36-41)
    out = _conv1d(cx, p + ".conv1", x, stride, 1)
    cx.taps[p + ".conv1"] = out
    out = _lrelu(_bn_apply(cx, p + ".bn1", out), SLOPE_BACKBONE, cx, p + ".a1")
    cx.taps[p + ".a1"] = out
    out = _conv1d(cx, p + ".conv2", out, 1, 1)
    out = _bn_apply(cx, p + ".bn2", out)
    if stride == 1:
        sc = x
    else:
        sc = _bn_apply(cx, p + ".shortcut.1", _conv1d(cx, p + ".shortcut.0", x, stride, 0))
    out = _lrelu(out + sc, SLOPE_BACKBONE, cx, p)
    cx.taps[p] = out
    return out


def _encoder(cx: _Ctx, p: str, x):
    # ResNet18Enc.forward (backbones.py:94-103)
    x = _conv1d(cx, p + ".conv1", x, 2, 1)
    cx.taps[p + ".conv1"] = x
    x = _lrelu(_bn_apply(cx, p + ".bn1", x), SLOPE_BACKBONE, cx, p + ".stem")
    cx.taps[p + ".stem"] = x
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        x = _block_enc(cx, f"{p}.layer{li}.0", x, stride)
        x = _block_enc(cx, f"{p}.layer{li}.1", x, 1)
    x = x.mean(dim=2)  # adaptive_avg_pool1d(x, 1).view(B, -1)
    return F.linear(x, cx.st[p + ".linear.weight"], cx.st[p + ".linear.bias"])


def _block_dec(cx: _Ctx, p: str, x, stride):
    # BasicBlockDec.forward (backbones.py:65-70); ResizeConv1d.forward (:13-16)
    out = _conv1d(cx, p + ".conv2", x, 1, 1)
    out = _lrelu(_bn_apply(cx, p + ".bn2", out), SLOPE_BACKBONE, cx, p + ".a2")
    cx.taps[p + ".a2"] = out
    if stride == 1:
        out = _bn_apply(cx, p + ".bn1", _conv1d(cx, p + ".conv1", out, 1, 1))
        sc = x
    else:
        out = _bn_apply(cx, p + ".bn1", _conv1d(cx, p + ".conv1.conv", _up_nearest(out, stride), 1, 1))
        sc = _bn_apply(cx, p + ".shortcut.1", _conv1d(cx, p + ".shortcut.0.conv", _up_nearest(x, stride), 1, 1))
    out = _lrelu(out + sc, SLOPE_BACKBONE, cx, p)
    cx.taps[p] = out
    return out


def _decoder(cx: _Ctx, p: str, x):
    # ResNet18Dec.forward (backbones.py:128-141)
    x = F.linear(x, cx.st[p + ".linear.weight"], cx.st[p + ".linear.bias"])
    x = _up_nearest(x.unsqueeze(-1), 4)
    cx.taps[p + ".linear"] = x
    for li, stride in ((4, 2), (3, 2), (2, 2), (1, 1)):
        x = _block_dec(cx, f"{p}.layer{li}.0", x, 1)
        x = _block_dec(cx, f"{p}.layer{li}.1", x, stride)
    x = _conv1d(cx, p + ".conv1.conv", _up_nearest(x, 2), 1, 1)
    x = x.reshape(x.shape[0], -1)
    cx.taps[p + ".conv1"] = x
    x = F.linear(x, cx.st[p + ".linear_out.weight"], cx.st[p + ".linear_out.bias"])
    return x.unsqueeze(1)


def _lin(cx, name, x):
    return F.linear(x, cx.st[name + ".weight"], cx.st[name + ".bias"])


def _decoder_fc(cx, p, z):
    # nn.Sequential(Linear, LeakyReLU(.2), Linear, BatchNorm1d, LeakyReLU(.2))  model.py:379-392
    z = _lrelu(_lin(cx, p + ".0", z), SLOPE_HEAD)
    z = _lin(cx, p + ".2", z)
    return _lrelu(_bn_apply(cx, p + ".3", z), SLOPE_HEAD)


def forward(st, cfg: CVAEConfig, x1, x2, src, cls=None, eps=None, train: bool = True, masks=None):
    """MultiModalCVAE.forward (model.py:424-432) / hippieUnimodalCVAE.forward (:62-72).

    x1: [B,1,L1]; x2: [B,1,L2] (None when unimodal); src/cls: int64 [B] (cls None ->
    class embedding := zeros_like(source_emb)); eps: the N(0,1) draw of `reparameterize`
    (model.py:397-400), injected so that CPU and GPU see the same noise.
    Returns (outputs dict, new_buffers dict, taps dict)."""
    cx = _Ctx(st, train, masks)
    source_emb = st["source_embedding.weight"][src]
    class_emb = st["class_embedding.weight"][cls] if cls is not None else torch.zeros_like(source_emb)
    if cfg.multimodal:
        h1 = _encoder(cx, "encoder_mod1", x1)
        h2 = _encoder(cx, "encoder_mod2", x2)
        cx.taps["h1"], cx.taps["h2"] = h1, h2
        h = torch.cat([h1, h2, source_emb, class_emb], dim=1)
        h = _lin(cx, "fusion_encoder.0", h)
        h = _lrelu(_bn_apply(cx, "fusion_encoder.1", h), SLOPE_HEAD)
        h = _lin(cx, "fusion_encoder.3", h)
    else:
        h = _encoder(cx, "encoder", x1)
        h = torch.cat([h, source_emb, class_emb], dim=1)
        h = _lin(cx, "encoder_fc.0", h)
        h = _lrelu(_bn_apply(cx, "encoder_fc.1", h), SLOPE_HEAD)
        h = _lin(cx, "encoder_fc.3", h)
        h = _lrelu(_bn_apply(cx, "encoder_fc.4", h), SLOPE_HEAD)
    mu = _lin(cx, "z_mean", h)
    logvar = _lin(cx, "z_log_var", h)
    std = torch.exp(0.5 * logvar)
    if eps is None:
        eps = torch.randn_like(std)
    z = mu + eps * std
    zc = torch.cat([z, source_emb, class_emb], dim=1)
    out = {"enc": h, "mu": mu, "logvar": logvar}
    if cfg.multimodal:
        out["dec1"] = _decoder(cx, "decoder_mod1", _decoder_fc(cx, "decoder_fc_mod1", zc))
        out["dec2"] = _decoder(cx, "decoder_mod2", _decoder_fc(cx, "decoder_fc_mod2", zc))
    else:
        out["dec1"] = _decoder(cx, "decoder", _decoder_fc(cx, "decoder_fc", zc))
    return out, cx.new_buffers, cx.taps


def loss_terms(out, x1, x2, beta, w1=1.0, w2=1.0, multimodal=True):
    """training_step arithmetic (model.py:465-474; unimodal :103-109).
    Returns (total, mse1, mse2, kl_mean)."""
    mse1 = ((x1 - out["dec1"]) ** 2).mean()
    kl = -0.5 * torch.sum(1 + out["logvar"] - out["mu"] ** 2 - torch.exp(out["logvar"]), dim=1)
    if multimodal:
        mse2 = ((x2 - out["dec2"]) ** 2).mean()
        mse = w1 * mse1 + w2 * mse2
    else:
        mse2 = torch.zeros_like(mse1)
        mse = mse1
    total = mse + beta * kl.mean()
    return total, mse1, mse2, kl.mean()


# --------------------------------------------------------------------------------------
# Optimiser side: clip_grad_norm_ + AdamW, restated element-wise
# --------------------------------------------------------------------------------------
def clip_coef(grads: Dict[str, torch.Tensor], max_norm: float):
    """torch.nn.utils.clip_grad_norm_ (norm_type 2): total = ||(||g_i||)_i||;
    coef = min(1, max_norm / (total + 1e-6))."""
    norms = torch.stack([g.norm(2) for g in grads.values()])
    total = norms.norm(2)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, coef


def adamw_update(p, g, m, v, step, lr, wd, b1=0.9, b2=0.999, eps=1e-8):
    """One torch.optim.AdamW element-wise update, in torch's single-tensor order
    (torch/optim/adam.py `_single_tensor_adam` with decoupled weight decay):
      p *= 1 - lr*wd; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g;
      denom = sqrt(v)/sqrt(1-b2^t) + eps; p -= (lr/(1-b1^t)) * m/denom."""
    p = p * (1 - lr * wd)
    m = m + (g - m) * (1 - b1)
    v = v * b2 + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def new_opt_state(st, cfg):
    return {"step": {n: 0 for n in param_names(cfg)},
            "exp_avg": {n: torch.zeros_like(st[n]) for n in param_names(cfg)},
            "exp_avg_sq": {n: torch.zeros_like(st[n]) for n in param_names(cfg)}}


def train_step(st, opt, cfg: CVAEConfig, x1, x2, labels, eps, *, lr, weight_decay, beta,
               w1=1.0, w2=1.0, max_norm: Optional[float] = 1.0, grad_hook=None, masks=None):
    """One Lightning-ordered optimisation step:
    training_step -> zero_grad -> backward -> clip_grad_norm_ -> AdamW.step.
    labels: int64 [B] (source only) or [B,2] = [class, source] (model.py:456-462).
    grad_hook(grads) may replace the gradients (used to emulate the DP all-reduce).
    Returns (new_state, new_opt, info) -- inputs are not mutated."""
    if labels.dim() == 2:
        cls, src = labels.unbind(1)
    else:
        cls, src = None, labels
    names = param_names(cfg)
    work = OrderedDict((k, (v.detach().clone().requires_grad_(True) if k in set(names) else v)) for k, v in st.items())
    out, new_buf, _ = forward(work, cfg, x1, x2, src, cls, eps, train=True, masks=masks)
    total, mse1, mse2, klm = loss_terms(out, x1, x2, beta, w1, w2, cfg.multimodal)
    total.backward()
    grads = {n: work[n].grad for n in names if work[n].grad is not None}
    if grad_hook is not None:
        grads = grad_hook(grads)
    info = {"loss": total.detach(), "mse1": mse1.detach(), "mse2": mse2.detach(), "kl": klm.detach(),
            "out": {k: v.detach() for k, v in out.items()},
            "grads_raw": {k: v.clone() for k, v in grads.items()}}
    if max_norm is not None:
        tn, coef = clip_coef(grads, max_norm)
        grads = {k: g * coef for k, g in grads.items()}
        info["grad_norm"] = tn
        info["clip_coef"] = coef
    new_st = OrderedDict((k, v.detach()) for k, v in st.items())
    new_st.update(new_buf)
    new_opt = {"step": dict(opt["step"]), "exp_avg": dict(opt["exp_avg"]), "exp_avg_sq": dict(opt["exp_avg_sq"])}
    for n, g in grads.items():  # params whose grad is None are skipped entirely, as torch does
        t = new_opt["step"][n] + 1
        p, m, v = adamw_update(st[n].detach(), g, opt["exp_avg"][n], opt["exp_avg_sq"][n], t, lr, weight_decay)
        new_st[n], new_opt["exp_avg"][n], new_opt["exp_avg_sq"][n], new_opt["step"][n] = p, m, v, t
    return new_st, new_opt, info


# --------------------------------------------------------------------------------------
# Data side (hippie/dataloading.py:27-56) and embedding post-processing
# --------------------------------------------------------------------------------------
def _fma(a, b, c):
    """fp32 fused multiply-add a*b+c with a single rounding (the product of two fp32 values
    is exact in fp64)."""
    return (a.double() * b.double() + c.double()).float()


def interp_linear(x: torch.Tensor, size: int) -> torch.Tensor:
    """F.interpolate(x[1,1,n], size=(size,), mode='linear', align_corners=False) on the
    last dim, as ATen's CPU kernel evaluates it (measured here: the build contracts both
    expressions into FMAs): scale = fp32(n)/size; src = max(fma(scale, i + 0.5, -0.5), 0);
    i0 = min(floor(src), n-1); i1 = min(i0+1, n-1); w1 = src - i0; w0 = 1 - w1;
    out = fma(x[i0], w0, x[i1] * w1)."""
    n = x.shape[-1]
    scale = (torch.tensor(float(n), dtype=torch.float32) / size)
    i = torch.arange(size, dtype=torch.float32)
    src = torch.clamp(_fma(scale, i + 0.5, torch.tensor(-0.5)), min=0.0)
    i0 = torch.clamp(src.floor().to(torch.long), max=n - 1)
    i1 = torch.clamp(i0 + 1, max=n - 1)
    w1 = torch.clamp(src - i0.to(torch.float32), 0.0, 1.0)
    w0 = 1.0 - w1
    return _fma(x[..., i0], w0, x[..., i1] * w1)


def dataset_item(wave_row, isi_row):
    """EphysDataset.__getitem__ with normalize=False (dataloading.py:27-50):
    fp32 cast; log(isi + 1); linear interpolation to 50 / 100; view(1, -1)."""
    w = torch.as_tensor(wave_row).float()
    t = torch.log(torch.as_tensor(isi_row).float() + 1)
    return interp_linear(w, 50).view(1, -1), interp_linear(t, 100).view(1, -1)


def zscore_rows(e: torch.Tensor, ddof: int) -> torch.Tensor:
    """Per-row z-score of the embedding: ddof=0 is get_embeddings_multimodal's np.std
    (scripts/train_model_with_multimodal.py:31); ddof=1 is get_embeddings' torch.std
    (scripts/utils.py:87-88)."""
    m = e.mean(dim=1, keepdim=True)
    n = e.shape[1]
    sd = torch.sqrt(((e - m) ** 2).sum(dim=1, keepdim=True) / (n - ddof))
    return (e - m) / sd


def synthetic_batch(n: int, seed: int = 1234, labelled: bool = False, pretrain_sources: bool = True):
    """SURVEY.md §8(d) synthetic inputs of the cellexplorer-celltype shape."""
    g = torch.Generator().manual_seed(seed)
    x1 = (0.365 * torch.randn(n, 1, 50, generator=g) + 0.019).clamp(-1, 1.3)
    x2 = torch.log1p(0.0157 * torch.randn(n, 1, 100, generator=g).abs())
    src = torch.randint(1, 5, (n,), generator=g) if pretrain_sources else torch.full((n,), 3)
    if labelled:
        cls = torch.randint(0, 4, (n,), generator=g)
        labels = torch.stack([cls, src], dim=1)
    else:
        labels = src
    eps_gen = torch.Generator().manual_seed(seed + 1)
    return x1, x2, labels, eps_gen
