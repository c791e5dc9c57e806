"""Lightning-free trainer for the engine-backed modules: the call order of `pl.Trainer.fit` that the reference's
scripts rely on (`pytorch_lightning` itself is not part of the reference and is not installed
here), `.ckpt` files in Lightning's dictionary layout, and the two callbacks the scripts use.

    trainer = Trainer(max_epochs=E, gradient_clip_val=1.0, callbacks=[ModelCheckpoint("val_loss"), EarlyStopping(...)])
    trainer.fit(module, train_loader, val_loader)
    module.load_state_dict(torch.load(ckpt.best_model_path)["state_dict"])     # scripts/...:706-707

Per `fit`: 2 sanity validation batches (eval mode; they draw reparameterisation noise like Lightning's do), then per
epoch  train(): training_step -> [gradient all-reduce] -> clip + AdamW  |  on_train_epoch_end  |  eval(): validation_step
per batch | on_validation_epoch_end -> epoch-mean `val_loss` -> ModelCheckpoint (top-1) and EarlyStopping.  Nothing in the
training loop synchronises with the host: losses stay on the device until the epoch mean is taken.

Data parallel: when `torch.distributed` is initialised every rank runs this loop on its own shard of the batches; the
flat gradient buffer is all-reduced inside `training_step`, overlapped with the encoders' backward pass
(hippie_b200/parallel.py:train_step_overlapped), the epoch-mean
validation loss is averaged over the ranks and rank 0 writes the checkpoints.
"""
from __future__ import annotations

import os
from typing import Iterable, Optional, Sequence

import torch
import torch.distributed as dist

CKPT_VERSION = "2.0.0+hippie_b200"  # value of the "pytorch-lightning_version" key (layout of Lightning >= 1.6)


class Callback:
    def on_fit_start(self, trainer, module):
        pass

    def on_validation_end(self, trainer, module, metrics):
        pass


class ModelCheckpoint(Callback):
    """pl.callbacks.ModelCheckpoint(monitor=..., save_top_k=1, mode=...) as the reference uses it
    (scripts/train_model_with_multimodal.py:682-684).  State lives on the object, so passing the same instance to a
    second Trainer keeps the best score of the first one -- exactly what happens in the reference's stage 2."""

    def __init__(self, monitor: str = "val_loss", save_top_k: int = 1, mode: str = "min", dirpath: Optional[str] = None,
                 filename: Optional[str] = None):
        assert mode in ("min", "max") and save_top_k in (0, 1), "the reference uses top-1 checkpoints"
        self.monitor, self.save_top_k, self.mode, self.dirpath, self.filename = monitor, save_top_k, mode, dirpath, filename
        self.best_model_path = ""
        self.best_model_score: Optional[float] = None

    def _better(self, cur: float) -> bool:
        if self.best_model_score is None:
            return True
        return cur < self.best_model_score if self.mode == "min" else cur > self.best_model_score

    def on_validation_end(self, trainer, module, metrics):
        cur = metrics.get(self.monitor)
        if cur is None or self.save_top_k == 0 or cur != cur or not self._better(cur):
            return
        dirpath = self.dirpath or os.path.join(trainer.log_dir, "checkpoints")
        name = self.filename or f"epoch={trainer.current_epoch}-step={trainer.global_step}"
        path = os.path.join(dirpath, name + ".ckpt")
        if trainer.is_global_zero:
            os.makedirs(dirpath, exist_ok=True)
            trainer.save_checkpoint(path)
            if self.best_model_path and self.best_model_path != path and os.path.exists(self.best_model_path):
                os.remove(self.best_model_path)
        self.best_model_path, self.best_model_score = path, float(cur)


class EarlyStopping(Callback):
    """pl.callbacks.EarlyStopping(monitor, patience, mode): stop when the monitored epoch value has not improved for
    `patience` consecutive validations (scripts/train_model_with_multimodal.py:685-687)."""

    def __init__(self, monitor: str = "val_loss", patience: int = 3, mode: str = "min", min_delta: float = 0.0):
        self.monitor, self.patience, self.mode, self.min_delta = monitor, patience, mode, abs(min_delta)
        self.best_score: Optional[float] = None
        self.wait_count = 0
        self.stopped_epoch = 0

    def on_validation_end(self, trainer, module, metrics):
        cur = metrics.get(self.monitor)
        if cur is None:
            return
        improved = self.best_score is None or (cur < self.best_score - self.min_delta if self.mode == "min"
                                               else cur > self.best_score + self.min_delta)
        if improved:
            self.best_score, self.wait_count = float(cur), 0
        else:
            self.wait_count += 1
            if self.wait_count >= self.patience:
                trainer.should_stop = True
                self.stopped_epoch = trainer.current_epoch


class LearningRateMonitor(Callback):
    """Accepted for signature compatibility (scripts/...:881); the learning rate is constant in the reference."""

    def __init__(self, logging_interval: str = "step"):
        self.logging_interval = logging_interval


def _limit(n_batches: Optional[int], limit) -> Optional[int]:
    """Lightning's limit_{train,val}_batches: None = all, float in [0, 1] = fraction, int = count."""
    if limit is None:
        return n_batches
    if isinstance(limit, float) and limit <= 1.0:
        return None if n_batches is None else int(n_batches * limit)
    return int(limit) if n_batches is None else min(int(limit), n_batches)


class Trainer:
    def __init__(self, max_epochs: int = 1, accelerator: Optional[str] = None, logger=None,
                 callbacks: Sequence[Callback] = (), limit_train_batches=None, limit_val_batches=None,
                 gradient_clip_val: Optional[float] = None, default_root_dir: str = "lightning_logs",
                 num_sanity_val_steps: int = 2, log_every_n_steps: int = 50, device: Optional[str] = None):
        if accelerator == "cpu":
            raise RuntimeError("hippie_b200 has no CPU path (accelerator='cpu'); it needs a CUDA device")
        self.max_epochs, self.logger, self.callbacks = max_epochs, logger, list(callbacks)
        self.limit_train_batches, self.limit_val_batches = limit_train_batches, limit_val_batches
        self.gradient_clip_val = gradient_clip_val
        self.num_sanity_val_steps, self.log_every_n_steps = num_sanity_val_steps, log_every_n_steps
        self.device = device
        self.current_epoch = 0
        self.global_step = 0
        self.should_stop = False
        self.callback_metrics = {}
        self._module = None
        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world_size > 1 else 0
        # lightning_logs/version_N like Lightning's default logger directory; rank 0 picks it (the other ranks would race
        # with rank 0 creating the directory) and everybody uses rank 0's choice
        n = 0
        while os.path.exists(os.path.join(default_root_dir, f"version_{n}")):
            n += 1
        if self.world_size > 1:
            box = [n]
            dist.broadcast_object_list(box, src=0)
            n = box[0]
        self.log_dir = os.path.join(default_root_dir, f"version_{n}")

    @property
    def is_global_zero(self) -> bool:
        return self.rank == 0

    # ---- checkpoints (Lightning's dictionary layout; torch.load(path)["state_dict"] is what the scripts read) ----
    def save_checkpoint(self, path: str):
        m = self._module
        ckpt = {"epoch": self.current_epoch, "global_step": self.global_step, "pytorch-lightning_version": CKPT_VERSION,
                "state_dict": {k: v.detach().cpu().clone() for k, v in m.state_dict().items()},
                "loops": {}, "callbacks": {}, "optimizer_states": [m.optimizer.state_dict()], "lr_schedulers": [],
                "hyper_parameters": {"learning_rate": m.lr, "weight_decay": m.weight_decay, "beta": m.beta}}
        tmp = path + ".tmp"
        torch.save(ckpt, tmp)
        os.replace(tmp, path)

    # ---- loops ---------------------------------------------------------------------------------------------------
    def _run_validation(self, module, loader, max_batches: Optional[int], sanity: bool):
        was_training = module.training
        module.eval()
        n = 0
        with torch.no_grad():
            for i, batch in enumerate(loader):
                if max_batches is not None and i >= max_batches:
                    break
                module.validation_step(batch, i)
                n += 1
        rows = float(sum(getattr(module, "_val_sizes", None) or [n]))  # before the epoch-end hook clears the sizes
        plain = module.on_validation_epoch_end() if n else float("nan")
        # what Lightning's callbacks monitor: the epoch aggregate of self.log("val_loss"), weighted by batch size
        mean = getattr(module, "val_loss_epoch", plain) if n else float("nan")
        if was_training:
            module.train()
        if self.world_size > 1:  # every rank takes part, whatever it saw: (sum of loss x rows, rows)
            ok = n > 0 and mean == mean
            t = torch.tensor([mean * rows if ok else 0.0, rows if ok else 0.0], dtype=torch.float64,
                             device=module.model._flat["params"].device)
            dist.all_reduce(t)
            n = int(float(t[1]) > 0)
            mean = float(t[0] / t[1]) if n else float("nan")
        if sanity or not n:
            return None
        return mean

    def wait_for_checkpoint(self):
        """Call before torch.load(ckpt.best_model_path) under data parallelism: rank 0 writes, everybody reads."""
        if self.world_size > 1:
            dist.barrier()

    def fit(self, module, train_dataloaders: Iterable, val_dataloaders: Optional[Iterable] = None):
        dev = self.device or (f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else None)
        if dev is None:
            raise RuntimeError("hippie_b200 needs a CUDA device to train; there is no CPU fallback")
        module.to(dev)
        self._module = module
        module.trainer = self
        module.world_size = self.world_size
        self.should_stop = False
        for cb in self.callbacks:
            cb.on_fit_start(self, module)
        n_train = len(train_dataloaders) if hasattr(train_dataloaders, "__len__") else None
        n_val = len(val_dataloaders) if val_dataloaders is not None and hasattr(val_dataloaders, "__len__") else None
        max_train, max_val = _limit(n_train, self.limit_train_batches), _limit(n_val, self.limit_val_batches)
        if val_dataloaders is not None and self.num_sanity_val_steps > 0:
            sanity = self.num_sanity_val_steps if max_val is None else min(self.num_sanity_val_steps, max_val)
            self._run_validation(module, val_dataloaders, sanity, sanity=True)
        clip = self.gradient_clip_val if self.gradient_clip_val else None
        for epoch in range(self.max_epochs):
            self.current_epoch = module.current_epoch = epoch
            module.train()
            for i, batch in enumerate(train_dataloaders):
                if max_train is not None and i >= max_train:
                    break
                module.training_step(batch, i)  # data parallel: all-reduces the gradients itself (overlapped)
                module.optimizer.step(max_norm=clip, grad_scale=module.grad_scale)
                self.global_step += 1
                if self.logger is not None and self.global_step % self.log_every_n_steps == 0 and self.is_global_zero:
                    self.logger.log_metrics({k: float(v) for k, v in module.logged.items()}, step=self.global_step)
            train_mean = module.on_train_epoch_end()
            module.model.check_device_flags()  # bad label indices / fp16 range: raised once per epoch (the epoch mean above
            metrics = {"train_loss_epoch": train_mean}  # has synchronised already)
            if val_dataloaders is not None:
                val_mean = self._run_validation(module, val_dataloaders, max_val, sanity=False)
                if val_mean is not None:
                    metrics["val_loss"] = val_mean
            self.callback_metrics = metrics
            if self.logger is not None and self.is_global_zero:
                self.logger.log_metrics({k: v for k, v in metrics.items() if v == v}, step=self.global_step)
            for cb in self.callbacks:
                cb.on_validation_end(self, module, metrics)
            if self.should_stop:
                break
        torch.cuda.synchronize()
        return self
