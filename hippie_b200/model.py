"""Drop-in mirror of the reference's `hippie/model.py` on top of the sm_100a engine.

Same class names, constructor signatures, `forward` / `training_step` / `validation_step` contracts and
`state_dict()` keys as the reference (hippie/model.py:12-72, 75-162, 350-432, 434-533); the arithmetic
is done by libhippie_b200.so.  Parameters and BatchNorm buffers are views into the engine's flat fp32
buffers, so `state_dict()` / `load_state_dict()` / `.ckpt` interchange keep working unchanged.

There is no CPU fallback: the classes can be constructed, saved and loaded on the CPU, but any
forward / training call requires the model to be on a CUDA device.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from ._base import _EngineModule, _Node  # noqa: F401
from . import backbones as _backbones  # noqa: F401  (registers the typed backbone containers of the module tree)
from .engine import Engine, LAYOUT_CONV_OKI  # noqa: F401
from .parallel import train_step_overlapped


class _CVAEModule(_EngineModule):
    """What the two cVAE classes share on top of the engine-backed base."""

    def reparameterize(self, mu, logvar):
        """reference hippie/model.py:397-400 (host-side convenience; the engine fuses this step)."""
        std = torch.exp(0.5 * logvar)
        return mu + torch.randn_like(std) * std

    def _draw_eps(self, B):
        # the reference draws randn_like(std) from the default generator of the tensor's device (hippie/model.py:399)
        return torch.randn(B, self.z_dim, device=self._flat["params"].device, dtype=torch.float32)

    def _src_cls(self, source_labels, class_labels):
        return (self._labels(source_labels, self.num_sources, "source label"),
                self._labels(class_labels, self.num_classes, "class label"))

    def _forward_impl(self, x1, x2, source_labels, class_labels, eps=None):
        self._require_cuda()
        eng = self._engine
        src, cls = self._src_cls(source_labels, class_labels)
        B = x1.shape[0]
        if eps is None:
            eps = self._draw_eps(B)
        if self.training:
            outs = eng.train_forward(x1, x2, src, cls, eps)
        else:
            outs = eng.eval_forward(x1, x2, src, cls, eps)
        return outs

    def _emb(self, t):
        B = t.shape[0]
        t = t.detach().to(self._flat["params"].device, torch.float32).contiguous()
        assert t.shape == (B, self.class_hidden_dim), f"embedding rows must be [B, {self.class_hidden_dim}]"
        return t


class MultiModalCVAE(_CVAEModule):
    """reference hippie/model.py:350-432."""

    def __init__(self, z_dim, output_size_wave, output_size_isi, class_hidden_dim, num_sources, num_classes,
                 max_batch=512):
        super().__init__(z_dim=z_dim, len_wave=output_size_wave, len_isi=output_size_isi,
                         class_hidden_dim=class_hidden_dim, num_sources=num_sources, num_classes=num_classes,
                         multimodal=True, max_batch=max_batch)
        self.output_size_wave, self.output_size_isi = output_size_wave, output_size_isi

    def forward(self, data1, data2, source_labels, class_labels=None, eps=None):
        x1, x2 = self._prep(data1, self.output_size_wave), self._prep(data2, self.output_size_isi)
        o = self._forward_impl(x1, x2, source_labels, class_labels, eps)
        return o["enc"], o["mu"], o["logvar"], o["dec1"], o["dec2"]

    def embed(self, data1, data2, source_labels, class_labels=None, zscore_ddof=-1):
        """Encoders + fusion only (get_embeddings_multimodal's `model(sample)[0]`, scripts/
        train_model_with_multimodal.py:22-34, without running the two decoders the reference discards)."""
        self._require_cuda()
        x1, x2 = self._prep(data1, self.output_size_wave), self._prep(data2, self.output_size_isi)
        src, cls = self._src_cls(source_labels, class_labels)
        return self._engine.embed(x1, x2, src, cls, zscore_ddof)

    @torch.no_grad()
    def encode(self, x1, x2, source_emb, class_emb):
        """reference hippie/model.py:402-408: (h, z_mean(h), z_log_var(h)) from the two inputs and the embedding ROWS.
        Forward only (no autograd graph): training goes through `training_step`."""
        self._require_cuda()
        a, b = self._prep(x1, self.output_size_wave), self._prep(x2, self.output_size_isi)
        return self._engine.encode(a, b, self._emb(source_emb), self._emb(class_emb), train=self.training)

    @torch.no_grad()
    def decode(self, z, source_emb, class_emb):
        """reference hippie/model.py:410-422: (recon1 [B,1,Lw], recon2 [B,1,Li]) from z and the embedding rows."""
        self._require_cuda()
        z = z.detach().to(self._flat["params"].device, torch.float32).contiguous()
        return self._engine.decode(z, self._emb(source_emb), self._emb(class_emb), train=self.training)


class hippieUnimodalCVAE(_CVAEModule):
    """reference hippie/model.py:12-72."""

    def __init__(self, z_dim, output_size, class_hidden_dim, num_sources, num_classes, max_batch=512):
        super().__init__(z_dim=z_dim, len_wave=output_size, len_isi=output_size, class_hidden_dim=class_hidden_dim,
                         num_sources=num_sources, num_classes=num_classes, multimodal=False, max_batch=max_batch)
        self.output_size = output_size

    def forward(self, data, source_labels, class_labels=None, eps=None):
        x = self._prep(data, self.output_size)
        o = self._forward_impl(x, None, source_labels, class_labels, eps)
        return o["enc"], o["mu"], o["logvar"], o["dec1"]

    def embed(self, data, source_labels, class_labels=None, zscore_ddof=-1):
        self._require_cuda()
        x = self._prep(data, self.output_size)
        src, cls = self._src_cls(source_labels, class_labels)
        return self._engine.embed(x, None, src, cls, zscore_ddof)

    @torch.no_grad()
    def encode(self, x, source_emb, class_emb):
        """reference hippie/model.py:50-56 (forward only)."""
        self._require_cuda()
        return self._engine.encode(self._prep(x, self.output_size), None, self._emb(source_emb), self._emb(class_emb),
                                   train=self.training)

    @torch.no_grad()
    def decode(self, z, source_emb, class_emb):
        """reference hippie/model.py:58-61 (forward only)."""
        self._require_cuda()
        z = z.detach().to(self._flat["params"].device, torch.float32).contiguous()
        return self._engine.decode(z, self._emb(source_emb), self._emb(class_emb), train=self.training)[0]


# ------------------------------------------------------------------------------------------------------
class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (reference hippie/model.py:447) fused with Lightning's gradient clipping
    (scripts/train_model_with_multimodal.py:55,701) over the engine's flat buffers: one reduction + one
    update launch instead of 285 per-tensor ops.  `state_dict()` keeps torch's AdamW format."""

    def __init__(self, model: _EngineModule, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self._model = model
        super().__init__(list(model.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                                        amsgrad=False, maximize=False, foreach=None, capturable=False,
                                                        differentiable=False, fused=None))
        self._step = 0
        self._step_cls = 0
        self.has_cls_grad = False
        self.last_scalars = None
        self._surgery_seen = model._surgery_count  # a class table assigned later is not this optimizer's parameter

    @torch.no_grad()
    def step(self, closure=None, max_norm: Optional[float] = None, grad_scale: float = 1.0):
        loss = closure() if closure is not None else None
        self._model._require_cuda()
        g = self.param_groups[0]
        self._step += 1
        has_cls = self.has_cls_grad and self._surgery_seen == self._model._surgery_count
        if has_cls:
            self._step_cls += 1
        self.last_scalars = self._model.engine.clip_adamw(
            g["lr"], g["weight_decay"], self._step, max_norm=max_norm, grad_scale=grad_scale, betas=g["betas"],
            eps=g["eps"], step_cls=max(self._step_cls, 1), has_cls_grad=has_cls)
        return loss

    def zero_grad(self, set_to_none: bool = True):
        pass  # the engine zero-fills the flat gradient buffer at the start of every train_fwd_bwd

    def state_dict(self):
        eng = self._model.engine
        state = {}
        if self._step > 0 and eng.exp_avg is not None:
            for i, p in enumerate(eng.params):
                is_cls = p.name == "class_embedding.weight"
                st = self._step_cls if is_cls else self._step
                if st == 0:
                    continue
                state[i] = {"step": torch.tensor(float(st)),
                            "exp_avg": Engine.view_of(eng.exp_avg, p).detach().clone(),
                            "exp_avg_sq": Engine.view_of(eng.exp_avg_sq, p).detach().clone()}
        groups = [{k: v for k, v in self.param_groups[0].items() if k != "params"}]
        groups[0]["params"] = list(range(len(eng.params)))
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        eng = self._model.engine
        self._model._require_cuda()
        g = sd["param_groups"][0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in g:
                self.param_groups[0][k] = g[k]
        self._step = self._step_cls = 0
        with torch.no_grad():
            eng.exp_avg.zero_(), eng.exp_avg_sq.zero_()
            for i, st in sd["state"].items():
                p = eng.params[int(i)]
                Engine.view_of(eng.exp_avg, p).copy_(st["exp_avg"])
                Engine.view_of(eng.exp_avg_sq, p).copy_(st["exp_avg_sq"])
                if p.name == "class_embedding.weight":
                    self._step_cls = int(st["step"])
                else:
                    self._step = max(self._step, int(st["step"]))


class _TrainModuleBase(nn.Module):
    """What the reference inherits from pl.LightningModule, reduced to what its scripts use."""
    _SCALAR_RING = 4096

    def __init__(self, base_model, alpha_max, learning_rate, weight_decay, beta):
        super().__init__()
        self.model = base_model
        self.lr = learning_rate
        self.weight_decay = weight_decay
        self.alpha_max = alpha_max
        self.beta = beta
        self.val_loss = []
        self.train_loss = []
        self.optimizer = FusedAdamW(self.model, lr=self.lr, weight_decay=self.weight_decay)
        self.logged = {}
        self.current_epoch = 0
        self.global_step = 0
        self._ring = None
        self._ring_pos = 0
        self._val_sizes = []
        self.val_loss_epoch = float("nan")
        self.world_size = 1  # set by the data-parallel trainer
        self.grad_scale = 1.0  # factor the optimizer applies to the (summed) gradients: 1 / world after an all-reduce

    def log(self, name, value, *a, **k):
        self.logged[name] = value

    def configure_optimizers(self):
        return self.optimizer

    def _scalars(self):
        dev = self.model._flat["params"].device
        if self._ring is None or self._ring.device != dev or self._ring_pos == self._SCALAR_RING:
            # a fresh block: the epoch lists (train_loss / val_loss) hold views into the earlier ones, which therefore
            # stay alive and are never overwritten, however many steps an epoch has
            self._ring = torch.zeros(self._SCALAR_RING, 8, dtype=torch.float32, device=dev)
            self._ring_pos = 0
        s = self._ring[self._ring_pos]
        self._ring_pos += 1
        return s

    @staticmethod
    def _mean(vals, weights=None):
        if not vals:
            return float("nan")
        if isinstance(vals[0], torch.Tensor):
            v = torch.stack([x.detach().double() for x in vals])  # one sync per epoch
            if weights:
                w = torch.tensor(weights, dtype=torch.float64, device=v.device)
                return ((v * w).sum() / w.sum()).item()
            return v.mean().item()
        return sum(vals) / len(vals)

    def on_validation_epoch_end(self):
        """The reference prints the plain mean of the per-batch losses (hippie/model.py:522-527); the value Lightning's
        ModelCheckpoint / EarlyStopping monitor is the epoch aggregate of `self.log("val_loss", ...)`, which weights
        every batch by its size.  Both are available: this returns the printed one, `val_loss_epoch` holds the other."""
        avg_loss = self._mean(self.val_loss)
        self.val_loss_epoch = self._mean(self.val_loss, self._val_sizes) if self._val_sizes else avg_loss
        print(f"Average validation loss is {avg_loss:.2f}")
        self.val_loss = []
        self._val_sizes = []
        return avg_loss

    def on_train_epoch_end(self):
        avg_loss = self._mean(self.train_loss)
        print(f"Average training loss is {avg_loss:.2f}")
        self.train_loss = []
        return avg_loss

    def _device_labels(self, labels):
        """(source, class | None) on the device; labels that arrive on the host are range-checked like nn.Embedding."""
        m = self.model
        src, cls = self._split_labels(labels)
        src = m._labels(src, m.num_sources, "source label")
        cls = m._labels(cls, m.num_classes, "class label")
        return src, cls

    @staticmethod
    def _split_labels(labels):
        # reference hippie/model.py:456-462: labels [B,2] = [class, source]; [B] = source only
        if labels.ndim == 2:
            class_labels, source_labels = labels.unbind(1)
            return source_labels, class_labels
        return labels, None


class MultiModalCVAETrainModule(_TrainModuleBase):
    """reference hippie/model.py:434-533."""

    def __init__(self, base_model, alpha_max=0.5, learning_rate=0.01, weight_decay=0.01, beta=1, mod1_weight=1.0,
                 mod2_weight=1.0):
        super().__init__(base_model, alpha_max, learning_rate, weight_decay, beta)
        self.mod1_weight = mod1_weight
        self.mod2_weight = mod2_weight

    def training_step(self, batch, batch_idx, eps=None):
        """forward + loss + backward in one fused engine call.  Returns the 0-dim total loss (device tensor);
        gradients are in the engine's flat buffer (exposed as `.grad` views).  Unlike the reference
        (`.item()` every step, hippie/model.py:480) nothing here synchronises with the host."""
        m = self.model
        m._require_cuda()
        data1, data2, labels = batch
        x1, x2 = m._prep(data1, m.output_size_wave), m._prep(data2, m.output_size_isi)
        src, cls = self._device_labels(labels)
        if eps is None:
            eps = m._draw_eps(x1.shape[0])
        s = self._scalars()
        # data parallel: the gradient all-reduce is issued from here so that it overlaps the encoders' backward pass
        self.grad_scale = train_step_overlapped(m.engine, x1, x2, src, cls, eps, float(self.beta), float(self.mod1_weight),
                                                float(self.mod2_weight), scalars=s)
        self.optimizer.has_cls_grad = cls is not None
        self.log("train_loss", s[0]), self.log("train_mse_loss1", s[1])
        self.log("train_mse_loss2", s[2]), self.log("train_kl_loss", s[3])
        self.train_loss.append(s[0])
        self.global_step += 1
        return s[0]

    def validation_step(self, batch, batch_idx, eps=None):
        m = self.model
        m._require_cuda()
        data1, data2, labels = batch
        x1, x2 = m._prep(data1, m.output_size_wave), m._prep(data2, m.output_size_isi)
        src, cls = self._device_labels(labels)
        if eps is None:
            eps = m._draw_eps(x1.shape[0])
        s = self._scalars()
        m.engine.eval_forward(x1, x2, src, cls, eps, float(self.beta), float(self.mod1_weight),
                              float(self.mod2_weight), scalars=s)
        self.val_loss.append(s[0]), self._val_sizes.append(int(x1.shape[0]))
        self.log("val_loss", s[0]), self.log("val_mse_loss1", s[1])
        self.log("val_mse_loss2", s[2]), self.log("val_kl_loss", s[3])
        return s[0]

    def forward(self, batch):
        data1, data2, labels = batch
        src, cls = self._split_labels(labels)
        return self.model(data1, data2, source_labels=src, class_labels=cls)


class hippieUnimodalEmbeddingModelCVAE(_TrainModuleBase):
    """reference hippie/model.py:75-162."""

    def __init__(self, base_model, alpha_max=0.5, learning_rate=0.01, weight_decay=0.01, beta=1):
        super().__init__(base_model, alpha_max, learning_rate, weight_decay, beta)

    def _step(self, batch, train, eps):
        m = self.model
        m._require_cuda()
        data, labels = batch
        x = m._prep(data, m.output_size)
        src, cls = self._device_labels(labels)
        if eps is None:
            eps = m._draw_eps(x.shape[0])
        s = self._scalars()
        if train:
            self.grad_scale = train_step_overlapped(m.engine, x, None, src, cls, eps, float(self.beta), 1.0, 1.0, scalars=s)
            self.optimizer.has_cls_grad = cls is not None
        else:
            m.engine.eval_forward(x, None, src, cls, eps, float(self.beta), 1.0, 1.0, scalars=s)
        return s

    def training_step(self, batch, batch_idx, eps=None):
        s = self._step(batch, True, eps)
        self.log("train_loss", s[0]), self.log("train_mse_loss", s[1]), self.log("train_kl_loss", s[3])
        self.train_loss.append(s[0])
        self.global_step += 1
        return s[0]

    def validation_step(self, batch, batch_idx, eps=None):
        s = self._step(batch, False, eps)
        self.val_loss.append(s[0]), self._val_sizes.append(int(batch[0].shape[0]))
        self.log("val_loss", s[0]), self.log("val_mse_loss", s[1]), self.log("val_kl_loss", s[3])
        return s[0]

    def forward(self, batch):
        data, labels = batch
        src, cls = self._split_labels(labels)
        return self.model(data, source_labels=src, class_labels=cls)
