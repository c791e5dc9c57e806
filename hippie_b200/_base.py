"""Base of the engine-backed nn.Modules (hippie_b200/model.py, hippie_b200/backbones.py): owns the Engine handle and the
flat buffers, reproduces the reference's module tree (hence its `state_dict()` keys) with parameters that are views
into the flat storage, and moves everything between the CPU (construction, save / load) and the GPU (compute)."""
from __future__ import annotations

import math
import re

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import Engine


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's module tree (and hence its state_dict keys)."""


class _EmbeddingNode(_Node):
    """`model.source_embedding(labels)` / `model.class_embedding(labels)` keep working (hippie/model.py:425-426): the
    rows are needed to call `encode` / `decode` the way the reference's `forward` does."""

    def forward(self, index):
        return F.embedding(index.to(self.weight.device), self.weight)


# (pattern on the dotted module name, node class) used when a container is created
NODE_TYPES = [(re.compile(r"(^|\.)(source|class)_embedding$"), _EmbeddingNode)]
# callables(root) run after the module tree is built: hippie_b200/backbones.py gives the backbone containers their types
TREE_HOOKS = []


def node_class(dotted: str):
    for pat, cls in NODE_TYPES:
        if pat.search(dotted):
            return cls
    return _Node


class _EngineModule(nn.Module):
    """Base of MultiModalCVAE / hippieUnimodalCVAE / ResNet18Enc / ResNet18Dec: owns the Engine and the flat buffers."""

    def __init__(self, *, z_dim, len_wave, len_isi, class_hidden_dim, num_sources, num_classes, multimodal,
                 max_batch=512):
        super().__init__()
        self.z_dim = z_dim
        self.class_hidden_dim = class_hidden_dim
        self.num_sources = num_sources
        self.num_classes = num_classes
        self._max_batch = max_batch
        self._cfg = dict(z_dim=z_dim, len_wave=len_wave, len_isi=len_isi, class_hidden_dim=class_hidden_dim,
                         num_sources=num_sources, num_classes=num_classes, multimodal=multimodal)
        self._surgery_count = 0  # bumped when class_embedding is replaced (optimizers built before do not train the new table)
        object.__setattr__(self, "_engine", Engine(max_batch=max_batch, **self._cfg))
        self._new_cpu_storage()
        self._param_objs = {}
        self._build_tree()
        self._reset_parameters()

    def _new_cpu_storage(self):
        eng = self._engine
        # CPU-resident flat storage until the module is moved to a CUDA device
        self._flat = {
            "params": torch.zeros(eng.param_floats),
            "bn_mean": torch.zeros(eng.bn_floats),
            "bn_var": torch.ones(eng.bn_floats),
            "bn_count": torch.zeros(len(eng.bns), dtype=torch.int64),
        }

    # ---- module tree ---------------------------------------------------------------------------------
    def _node(self, dotted: str) -> nn.Module:
        m = self
        path = []
        for part in dotted.split("."):
            path.append(part)
            if part not in m._modules:
                nn.Module.__setattr__(m, part, node_class(".".join(path))())
            m = m._modules[part]
        return m

    def _build_tree(self):
        eng = self._engine
        bn_by_name = {b.name: b for b in eng.bns}
        for p in eng.params:
            mod_name, leaf = p.name.rsplit(".", 1)
            node = self._node(mod_name)
            param = nn.Parameter(Engine.view_of(self._flat["params"], p))
            node.register_parameter(leaf, param)
            self._param_objs[p.name] = param
            if leaf == "bias" and mod_name in bn_by_name:  # BatchNorm: buffers follow weight, bias
                b = bn_by_name[mod_name]
                node.register_buffer("running_mean", self._flat["bn_mean"][b.offset:b.offset + b.channels])
                node.register_buffer("running_var", self._flat["bn_var"][b.offset:b.offset + b.channels])
                node.register_buffer("num_batches_tracked", self._flat["bn_count"][b.index])
        for hook in TREE_HOOKS:
            hook(self)

    def _rebind(self):
        """Points every nn.Parameter / buffer at the current flat storage (after a device move)."""
        eng = self._engine
        for p in eng.params:
            param = self._param_objs[p.name]
            param.data = Engine.view_of(self._flat["params"], p)
            param.grad = None
        for b in eng.bns:
            node = self._node(b.name)
            node._buffers["running_mean"] = self._flat["bn_mean"][b.offset:b.offset + b.channels]
            node._buffers["running_var"] = self._flat["bn_var"][b.offset:b.offset + b.channels]
            node._buffers["num_batches_tracked"] = self._flat["bn_count"][b.index]

    def _attach_grads(self):
        """Exposes the engine's flat gradient buffer as `.grad` views (what loss.backward() fills in torch)."""
        eng = self._engine
        if eng.flat_grads is None:
            return
        for p in eng.params:
            self._param_objs[p.name].grad = Engine.view_of(eng.flat_grads, p)

    def _reset_parameters(self):
        """nn.Conv1d / nn.Linear: kaiming_uniform_(a=sqrt(5)) weights, U(+-1/sqrt(fan_in)) biases;
        nn.Embedding: N(0,1); nn.BatchNorm1d: ones / zeros -- drawn in construction order on the CPU
        generator, as the reference does when it builds the model right after torch.manual_seed(42)."""
        eng = self._engine
        bn_names = {b.name for b in eng.bns}
        fan_in_of = {}
        with torch.no_grad():
            for p in eng.params:
                mod_name, leaf = p.name.rsplit(".", 1)
                view = Engine.view_of(self._flat["params"], p)
                if mod_name in bn_names:
                    view.fill_(1.0 if leaf == "weight" else 0.0)
                elif mod_name.endswith("_embedding"):
                    view.copy_(torch.empty(p.shape).normal_())
                elif leaf == "weight":
                    fan_in = p.shape[1] * (p.shape[2] if len(p.shape) == 3 else 1)
                    fan_in_of[mod_name] = fan_in
                    gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
                    bound = math.sqrt(3.0) * gain / math.sqrt(fan_in)
                    view.copy_(torch.empty(p.shape).uniform_(-bound, bound))
                else:
                    bound = 1 / math.sqrt(fan_in_of[mod_name])
                    view.copy_(torch.empty(p.shape).uniform_(-bound, bound))

    # ---- device moves ----------------------------------------------------------------------------------
    def _apply(self, fn, recurse=True):
        probe = fn(torch.zeros(1, dtype=torch.float32, device=self._flat["params"].device))
        if probe.dtype != torch.float32:
            raise TypeError("hippie_b200 models are fp32 only (the reference trains in fp32)")
        if probe.device == self._flat["params"].device:
            return self
        eng = self._engine
        if probe.device.type == "cuda":
            old = self._flat
            eng.allocate(probe.device)
            eng.flat_params.copy_(old["params"])
            eng.bn_mean.copy_(old["bn_mean"])
            eng.bn_var.copy_(old["bn_var"])
            eng.bn_count.copy_(old["bn_count"])
            self._flat = {"params": eng.flat_params, "bn_mean": eng.bn_mean, "bn_var": eng.bn_var,
                          "bn_count": eng.bn_count}
        else:
            self._flat = {k: v.detach().to(probe.device).clone() for k, v in self._flat.items()}
        self._rebind()
        return self

    # ---- engine rebuilds (another table size, input length or batch capacity) ---------------------------
    def _rebuild_engine(self, keep_state: bool = True, skip=(), **cfg_changes):
        """Creates a new engine for a changed configuration and carries the named state (parameters, BatchNorm buffers,
        AdamW moments) over by name; tensors whose name is in `skip` or whose shape changed keep their fresh values."""
        old_eng, dev = self._engine, self._flat["params"].device
        state = {k: v.detach().clone() for k, v in self.state_dict().items()} if keep_state else {}
        moments = {}
        if keep_state and old_eng.exp_avg is not None:
            for p in old_eng.params:
                moments[p.name] = (p.shape, Engine.view_of(old_eng.exp_avg, p).clone(),
                                   Engine.view_of(old_eng.exp_avg_sq, p).clone())
        max_batch = cfg_changes.pop("max_batch", self._max_batch)
        self._cfg.update(cfg_changes)
        self._max_batch = max_batch
        object.__setattr__(self, "_engine", Engine(max_batch=max_batch, **self._cfg))
        del old_eng
        for name in list(self._modules):
            del self._modules[name]
        self._new_cpu_storage()
        self._param_objs = {}
        self._build_tree()
        eng = self._engine
        own = nn.Module.state_dict(self)
        with torch.no_grad():
            for k, v in state.items():
                if k in own and k not in skip and own[k].shape == v.shape:
                    own[k].copy_(v)
        if dev.type == "cuda":
            self._apply(lambda t: t.to(dev))
            if eng.exp_avg is not None:
                with torch.no_grad():
                    for p in eng.params:
                        if p.name in moments and p.name not in skip and moments[p.name][0] == p.shape:
                            Engine.view_of(eng.exp_avg, p).copy_(moments[p.name][1])
                            Engine.view_of(eng.exp_avg_sq, p).copy_(moments[p.name][2])
        return self

    def __setattr__(self, name, value):
        # `model.class_embedding = nn.Embedding(k, class_hidden_dim)`: the reference's stage-3 surgery
        # (scripts/train_model_with_multimodal.py:378-379, scripts/train_model.py:299-300)
        if name == "class_embedding" and isinstance(value, nn.Module) and "_engine" in self.__dict__ \
                and self._cfg["multimodal"] in (0, 1, False, True):
            return self._replace_class_embedding(value)
        return super().__setattr__(name, value)

    def _replace_class_embedding(self, emb: nn.Module):
        w = getattr(emb, "weight", None)
        if w is None or w.dim() != 2 or w.shape[1] != self.class_hidden_dim:
            raise ValueError(f"class_embedding must be an nn.Embedding(k, {self.class_hidden_dim})")
        k = int(w.shape[0])
        self._rebuild_engine(skip=("class_embedding.weight",), num_classes=k)
        self.num_classes = k
        with torch.no_grad():
            self._param_objs["class_embedding.weight"].copy_(w.detach().to(self._flat["params"].device, torch.float32))
        # As in the reference, an optimizer built BEFORE the assignment does not know the new table (torch's AdamW holds
        # the old Parameter): FusedAdamW compares this counter with the one it saw at construction.
        self._surgery_count += 1

    @property
    def engine(self) -> Engine:
        return self._engine

    @property
    def on_cuda(self) -> bool:
        return self._flat["params"].is_cuda

    def _require_cuda(self):
        if not self.on_cuda:
            raise RuntimeError("hippie_b200 has no CPU path: move the model to a CUDA device (model.to('cuda'))")

    def _prep(self, data, n):
        dev = self._flat["params"].device
        data = data.to(dev, torch.float32, non_blocking=True).contiguous()
        if data.shape[0] > self._max_batch:
            raise ValueError(f"batch {data.shape[0]} exceeds max_batch={self._max_batch} this model was built with")
        assert data.numel() == data.shape[0] * n, f"expected [B,1,{n}] input, got {tuple(data.shape)}"
        return data

    def _labels(self, t, bound=None, what="label"):
        """Moves a label tensor to the device.  Labels that arrive as CPU tensors are range-checked here, like
        nn.Embedding does (IndexError); device tensors are checked by the kernels (Engine.raise_on_flags)."""
        if t is None:
            return None
        if bound is not None and not t.is_cuda and t.numel():
            lo, hi = int(t.min()), int(t.max())
            if lo < 0 or hi >= bound:
                raise IndexError(f"index out of range in self: {what} {lo if lo < 0 else hi} outside [0, {bound})")
        return t.to(self._flat["params"].device, torch.int64, non_blocking=True).contiguous()

    def check_device_flags(self):
        """Raises what the kernels flagged since the last call (bad label indices, fp16 pair-plane saturation).
        Synchronises the stream: the trainer calls it once per epoch."""
        if self.on_cuda:
            self._engine.raise_on_flags()
