"""Drop-in mirror of the reference's `hippie/backbones.py` on top of the sm_100a engine.

`ResNet18Enc` / `ResNet18Dec` (reference hippie/backbones.py:73-103, :106-141) keep their constructor signatures,
`forward` contracts, module tree and `state_dict()` keys; the arithmetic is the engine's layer program
(`hippie_encoder_forward` / `hippie_decoder_forward`, include/hippie_b200.h).  As modules of their own they are
forward-only (`.train()`: batch statistics + running-statistics update, `.eval()`: running statistics); training runs
through the cVAE classes of hippie_b200/model.py, whose engines execute the same layer program with its backward pass.

`BasicBlockEnc`, `BasicBlockDec`, `ResizeConv1d` (reference :19-41, :44-70, :6-16) are the typed containers of that tree
(`enc.layer2[0].conv1.weight`, `dec.layer4[1].shortcut[0].conv.weight`, ...): they carry the parameters under the
reference's names, but a block never runs on its own -- the engine executes whole backbones.

There is no CPU fallback: construction, `state_dict()` / `load_state_dict()` work on the CPU, `forward` needs CUDA.
"""
from __future__ import annotations

import re

import torch

from ._base import TREE_HOOKS, _EmbeddingNode, _EngineModule, _Node
from .engine import KIND_DECODER, KIND_ENCODER


class _BlockNode(_Node):
    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} is a parameter container of the engine-backed module tree; the layer "
                           "program runs whole backbones (call the parent ResNet18Enc / ResNet18Dec or the cVAE model)")

    # nn.Sequential-style access used by code that walks the reference tree: layer1[0], shortcut[1], len(layer1)
    def __getitem__(self, i):
        return self._modules[str(i)]

    def __len__(self):
        return len(self._modules)


class ResizeConv1d(_BlockNode):
    """reference hippie/backbones.py:6-16 (nearest x`scale_factor` + Conv1d k3 p1 with bias); parameters `conv.weight`,
    `conv.bias`."""


class BasicBlockEnc(_BlockNode):
    """reference hippie/backbones.py:19-41; parameters conv1, bn1, conv2, bn2 (+ shortcut.0, shortcut.1 when stride 2)."""


class BasicBlockDec(_BlockNode):
    """reference hippie/backbones.py:44-70; parameters conv2, bn2, conv1 (a ResizeConv1d when stride 2), bn1
    (+ shortcut.0 = ResizeConv1d, shortcut.1 when stride 2)."""


class _Sequential(_BlockNode):
    """The nn.Sequential containers of the reference tree (layerN, shortcut)."""


def _type_tree(root):
    """Gives the containers of the tree the reference's types: blocks by position (`layerN.M`; encoder or decoder by the
    name of the backbone they sit in, or by the class of a stand-alone backbone), ResizeConv1d by structure (a container
    whose only child is `conv`)."""
    for name, m in root.named_modules():
        if not isinstance(m, _Node) or isinstance(m, _EmbeddingNode) or m is root:
            continue
        if re.search(r"(^|\.)layer\d\.\d$", name):
            enc = name.split(".")[0].startswith("encoder") or type(root).__name__ == "ResNet18Enc"
            m.__class__ = BasicBlockEnc if enc else BasicBlockDec
        elif re.search(r"(^|\.)layer\d$|\.shortcut$", name):
            m.__class__ = _Sequential
        elif "conv" in m._modules and not m._parameters:
            m.__class__ = ResizeConv1d


TREE_HOOKS.append(_type_tree)


class ResNet18Enc(_EngineModule):
    """reference hippie/backbones.py:73-103: x [B, nc=1, L] -> [B, 2 * z_dim]."""

    def __init__(self, num_blocks=[2, 2, 2, 2], z_dim=10, nc=1, input_size=50, max_batch=512):
        if list(num_blocks) != [2, 2, 2, 2] or nc != 1:
            raise ValueError("the engine implements the ResNet-18 layout of the reference (num_blocks [2,2,2,2], nc=1)")
        super().__init__(z_dim=z_dim, len_wave=input_size, len_isi=input_size, class_hidden_dim=1, num_sources=1,
                         num_classes=1, multimodal=KIND_ENCODER, max_batch=max_batch)
        self.in_planes = 512  # value the reference's constructor leaves behind (:76-83)

    @torch.no_grad()
    def forward(self, x):
        self._require_cuda()
        L, B = x.shape[-1], x.shape[0]
        if L != self._cfg["len_wave"] or B > self._max_batch:  # the reference module takes any length / batch
            self._rebuild_engine(len_wave=L, len_isi=L, max_batch=max(B, self._max_batch))
        return self._engine.encoder_forward(self._prep(x, L), 0, train=self.training)


class ResNet18Dec(_EngineModule):
    """reference hippie/backbones.py:106-141: x [B, 2 * z_dim] -> [B, nc=1, output_size]."""

    def __init__(self, output_size: int = 64, num_blocks=[2, 2, 2, 2], z_dim=10, nc=1, max_batch=512):
        if list(num_blocks) != [2, 2, 2, 2] or nc != 1:
            raise ValueError("the engine implements the ResNet-18 layout of the reference (num_blocks [2,2,2,2], nc=1)")
        super().__init__(z_dim=z_dim, len_wave=output_size, len_isi=output_size, class_hidden_dim=1, num_sources=1,
                         num_classes=1, multimodal=KIND_DECODER, max_batch=max_batch)
        self.in_planes = 64
        self.output_size = output_size

    @torch.no_grad()
    def forward(self, x):
        self._require_cuda()
        B = x.shape[0]
        if B > self._max_batch:
            self._rebuild_engine(max_batch=B)
        d = x.detach().to(self._flat["params"].device, torch.float32).contiguous()
        return self._engine.decoder_forward(d, 0, train=self.training)


def test_decoder(device="cuda"):
    """The reference's own test (hippie/backbones.py:156-165), on the engine."""
    sample = torch.randn(8, 20)
    model = ResNet18Dec(output_size=50).to(device)
    output = model(sample)
    assert output.shape == (8, 1, 50)

    sample = torch.randn(8, 20)
    model = ResNet18Dec(output_size=100).to(device)
    output = model(sample)
    assert output.shape == (8, 1, 100)
