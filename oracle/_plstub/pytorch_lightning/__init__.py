"""Minimal stand-in for `pytorch_lightning` (not installed; no network), used ONLY by
oracle/make_golden.py to import /root/reference/hippie/model.py in the build container.
It supplies the names that module touches at import / step time and nothing else."""
import types

import torch.nn as nn


class LightningModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.trainer = types.SimpleNamespace(max_epochs=1)
        self.current_epoch = 0
        self.logged = {}

    def log(self, name, value, *a, **k):
        self.logged[name] = value
