"""Host-side datasets and sampler with the reference's semantics (reference hippie/dataloading.py:18-151).

Per item: float32 cast, log(isi + 1), linear interpolation (align_corners=False) of the waveform to 50 and of
the ISI histogram to 100 samples, `view(1, -1)`; labels as int64.  Indexing, shuffling and label encoding stay
on the host so they are bit-identical to the reference; the tensors they produce are what the engine consumes.

Differences from the reference, all additive:
  * `EphysDatasetLabeled` accepts mode="both" (the reference's multimodal CLI passes it,
    scripts/train_model_with_multimodal.py:638, but the reference class asserts it away -- SURVEY.md F4);
  * `normalize=True` works (the reference calls np.min on a tensor and raises, dataloading.py:84);
  * `BalancedBatchSampler` takes an optional `seed` for its over-sampling draws (the reference uses the
    unseeded global `random`); seed=None keeps the reference behaviour;
  * `EphysTensorDataset.batch()` serves whole pre-transformed batches (one gather instead of B __getitem__ calls).
"""
from __future__ import annotations

import random
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils.data import Dataset, Sampler

WAVE_LEN = 50
ISI_LEN = 100


def _transform(waveform: torch.Tensor, isi_dist: torch.Tensor, normalize: bool):
    waveform = waveform.float()
    isi_dist = torch.log(isi_dist.float() + 1)
    if normalize:
        lo, hi = waveform.min(), waveform.max()
        waveform = (waveform - lo) / (hi - lo) * 2 - 1
        isi_dist = (isi_dist - isi_dist.mean()) / isi_dist.std()
    waveform = F.interpolate(waveform.view(1, 1, -1), size=(WAVE_LEN,), mode="linear").view(1, -1)
    isi_dist = F.interpolate(isi_dist.view(1, 1, -1), size=(ISI_LEN,), mode="linear").view(1, -1)
    return waveform, isi_dist


class EphysDataset(Dataset):
    """reference hippie/dataloading.py:18-59."""

    def __init__(self, waveforms, isi_dists, mode, normalize=True):
        self.waveforms = np.array(waveforms)
        self.isi_dists = np.array(isi_dists)
        assert mode in ("wave", "time", "both")
        self.mode = mode
        assert len(self.waveforms) == len(self.isi_dists)
        self.normalize = normalize

    def __getitem__(self, idx):
        waveform, isi_dist = _transform(torch.as_tensor(self.waveforms[idx, ...]),
                                        torch.as_tensor(self.isi_dists[idx, ...]), self.normalize)
        if self.mode == "wave":
            return waveform, -1
        if self.mode == "time":
            return isi_dist, -1
        return waveform, isi_dist

    def __len__(self):
        return len(self.waveforms)


class EphysDatasetLabeled(Dataset):
    """reference hippie/dataloading.py:62-104, plus the missing mode="both" -> (wave, isi, label)."""

    def __init__(self, waveforms, isi_dists, labels, mode, normalize=True):
        self.waveforms = np.array(waveforms)
        self.isi_dists = np.array(isi_dists)
        self.labels = np.array(labels)
        assert mode in ("wave", "time", "both")
        self.mode = mode
        assert len(self.waveforms) == len(self.isi_dists)
        assert len(self.waveforms) == len(self.labels)
        self.normalize = normalize

    def __getitem__(self, idx):
        waveform, isi_dist = _transform(torch.as_tensor(self.waveforms[idx, ...]),
                                        torch.as_tensor(self.isi_dists[idx, ...]), self.normalize)
        label = torch.as_tensor(self.labels[idx]).long()
        if self.mode == "wave":
            return waveform, label
        if self.mode == "time":
            return isi_dist, label
        return waveform, isi_dist, label

    def __len__(self):
        return len(self.waveforms)


class EphysTensorDataset(Dataset):
    """The same transform applied once to every row up front (rows may have different raw widths per source, so
    the transform is per source table), stored as dense [N,1,50] / [N,1,100] / labels tensors.  `batch(indices)`
    gathers a whole batch in one call -- the Python per-item path of the reference tops out near 20 K items/s."""

    def __init__(self, waveforms, isi_dists, labels=None, normalize=False):
        wf, isi = np.asarray(waveforms), np.asarray(isi_dists)
        assert len(wf) == len(isi)
        w = torch.as_tensor(wf).float()
        t = torch.log(torch.as_tensor(isi).float() + 1)
        if normalize:
            lo, hi = w.min(dim=1, keepdim=True).values, w.max(dim=1, keepdim=True).values
            w = (w - lo) / (hi - lo) * 2 - 1
            t = (t - t.mean(dim=1, keepdim=True)) / t.std(dim=1, keepdim=True)
        self.wave = F.interpolate(w.unsqueeze(1), size=(WAVE_LEN,), mode="linear")
        self.isi = F.interpolate(t.unsqueeze(1), size=(ISI_LEN,), mode="linear")
        self.labels = None if labels is None else torch.as_tensor(np.asarray(labels)).long()

    @classmethod
    def concat(cls, parts):
        """torch.utils.data.ConcatDataset over already-transformed tables (their raw widths may differ)."""
        out = cls.__new__(cls)
        out.wave = torch.cat([p.wave for p in parts])
        out.isi = torch.cat([p.isi for p in parts])
        labelled = [p.labels is not None for p in parts]
        assert all(labelled) or not any(labelled), "either every table carries labels or none does"
        out.labels = torch.cat([p.labels for p in parts]) if all(labelled) and parts else None
        return out

    def __len__(self):
        return self.wave.shape[0]

    def __getitem__(self, idx):
        if self.labels is None:
            return self.wave[idx], self.isi[idx]
        return self.wave[idx], self.isi[idx], self.labels[idx]

    def batch(self, indices):
        idx = torch.as_tensor(indices, dtype=torch.long)
        if self.labels is None:
            return self.wave[idx], self.isi[idx]
        return self.wave[idx], self.isi[idx], self.labels[idx]


class BalancedBatchSampler(Sampler):
    """Round-robin over classes with the minority classes over-sampled to the majority count
    (reference hippie/dataloading.py:107-151)."""

    def __init__(self, dataset, labels=None, seed: Optional[int] = None):
        if labels is None:
            raise Exception("You should pass the tensor of labels to the constructor as second argument")
        self.labels = labels
        rng = random if seed is None else random.Random(seed)
        self.dataset = {}
        for idx in range(len(dataset)):
            self.dataset.setdefault(self._get_label(dataset, idx), []).append(idx)
        self.balanced_max = max(len(v) for v in self.dataset.values()) if self.dataset else 0
        for label in self.dataset:
            while len(self.dataset[label]) < self.balanced_max:
                self.dataset[label].append(rng.choice(self.dataset[label]))
        self.keys = list(self.dataset.keys())
        self.currentkey = 0
        self.indices = [-1] * len(self.keys)

    def _get_label(self, dataset, idx, labels=None):
        return self.labels[idx].item()

    def __iter__(self):
        while self.indices[self.currentkey] < self.balanced_max - 1:
            self.indices[self.currentkey] += 1
            yield self.dataset[self.keys[self.currentkey]][self.indices[self.currentkey]]
            self.currentkey = (self.currentkey + 1) % len(self.keys)
        self.indices = [-1] * len(self.keys)

    def __len__(self):
        return self.balanced_max * len(self.keys)


class EphysBatchLoader:
    """Whole-batch loader over an `EphysTensorDataset`: the index stream, the batch boundaries and the consumption of
    torch's default generator are those of `torch.utils.data.DataLoader(dataset, batch_size, shuffle=..., sampler=...)`
    (one int64 for the iterator's base seed, one more to seed RandomSampler's private generator when shuffling), but a
    batch is ONE gather from the pre-transformed tensors instead of B `__getitem__` calls + collation -- the reference's
    per-item path tops out near 20 K items/s, a B200 consumes > 130 K samples/s.

    `indices` restricts the loader to a subset (what `torch.utils.data.Subset(dataset, indices)` does); `rank` / `world`
    give every data-parallel rank the r-th contiguous slice of each global batch of `world * batch_size` indices
    (hippie_b200/parallel.py:shard_batch_indices).  With `pin_memory=True` batches are staged in pinned host memory so
    that the module's `.to(device, non_blocking=True)` copies overlap the previous step."""

    def __init__(self, dataset: "EphysTensorDataset", batch_size: int, shuffle: bool = False, sampler=None,
                 indices=None, drop_last: bool = False, pin_memory: bool = False, rank: int = 0, world: int = 1):
        assert not (shuffle and sampler is not None), "sampler option is mutually exclusive with shuffle"
        self.dataset, self.batch_size, self.shuffle, self.sampler = dataset, int(batch_size), shuffle, sampler
        self.indices = None if indices is None else [int(i) for i in indices]
        self.drop_last, self.pin_memory, self.rank, self.world = drop_last, pin_memory, rank, world

    def _n(self) -> int:
        if self.sampler is not None:
            return len(self.sampler)
        return len(self.dataset) if self.indices is None else len(self.indices)

    def _tail_kept(self, tail: int) -> bool:
        """The ragged last global batch: dropped with `drop_last`; under data parallelism also when it cannot give every
        rank at least two rows (training-mode BatchNorm needs two, and a rank that skipped the step would leave the others
        waiting in the gradient all-reduce) -- every rank takes the same decision, so the step counts stay equal."""
        if tail == 0 or self.drop_last:
            return False
        return self.world == 1 or tail >= 2 * self.world

    def __len__(self) -> int:
        per_step = self.batch_size * self.world
        n = self._n()
        return n // per_step + (1 if self._tail_kept(n % per_step) else 0)

    def _order(self):
        torch.empty((), dtype=torch.int64).random_()  # DataLoader: _base_seed of the iterator
        n = self._n()
        if self.sampler is not None:
            order = [int(i) for i in self.sampler]
        elif self.shuffle:  # RandomSampler without replacement, private generator seeded from the default one
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(n, generator=g).tolist()
        else:
            order = list(range(n))
        return order if self.indices is None else [self.indices[i] for i in order]

    def __iter__(self):
        order = self._order()
        per_step = self.batch_size * self.world
        for lo in range(0, len(order), per_step):
            chunk = order[lo:lo + per_step]
            if len(chunk) < per_step:
                if not self._tail_kept(len(chunk)):
                    break
                # the tail is split evenly (contiguous shares in rank order), not by batch_size: no rank is left empty
                base, rem = divmod(len(chunk), self.world)
                a = self.rank * base + min(self.rank, rem)
                mine = chunk[a:a + base + (1 if self.rank < rem else 0)]
            else:
                mine = chunk[self.rank * self.batch_size:(self.rank + 1) * self.batch_size]
            batch = self.dataset.batch(mine)
            if self.pin_memory and torch.cuda.is_available():
                batch = tuple(t if t.is_cuda else t.pin_memory() for t in batch)  # DeviceTable batches are on the GPU already
            yield batch


class DeviceTable:
    """A raw source table (waveforms + ISI histograms as pandas reads them, float64) resident on the GPU; `batch(indices)`
    produces the model inputs of those rows with ONE kernel launch per modality (hippie_preprocess_batch): float32 cast,
    log(isi + 1), linear interpolation to 50 / 100 samples -- what `EphysDataset.__getitem__` does per item on the host
    (reference hippie/dataloading.py:27-56).  Waveforms are bit-exact with the host path, ISI values within one ulp before
    interpolation.  Indices come from the unchanged host samplers (`EphysBatchLoader._order`, `BalancedBatchSampler`)."""

    def __init__(self, waveforms, isi_dists, labels=None, device="cuda"):
        from . import _lib
        self._L = _lib.lib()
        self.device = torch.device(device)
        self.wave = torch.as_tensor(np.asarray(waveforms, dtype=np.float64)).contiguous().to(self.device)
        self.isi = torch.as_tensor(np.asarray(isi_dists, dtype=np.float64)).contiguous().to(self.device)
        assert self.wave.shape[0] == self.isi.shape[0]
        self.labels = None if labels is None else torch.as_tensor(np.asarray(labels)).long().to(self.device)

    def __len__(self):
        return self.wave.shape[0]

    def batch(self, indices):
        import ctypes as C
        idx = torch.as_tensor(indices, dtype=torch.int64).to(self.device, non_blocking=True).contiguous()
        B = idx.numel()
        x1 = torch.empty(B, 1, WAVE_LEN, dtype=torch.float32, device=self.device)
        x2 = torch.empty(B, 1, ISI_LEN, dtype=torch.float32, device=self.device)
        ptr = lambda t: C.c_void_p(t.data_ptr())
        rc = self._L.hippie_preprocess_batch(ptr(self.wave), self.wave.shape[1], ptr(self.isi), self.isi.shape[1], ptr(idx), B,
                                             ptr(x1), WAVE_LEN, ptr(x2), ISI_LEN,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(f"hippie_preprocess_batch failed ({rc})")
        if self.labels is None:
            return x1, x2
        return x1, x2, self.labels[idx]

    @classmethod
    def concat(cls, parts):
        """torch.utils.data.ConcatDataset over device tables whose raw widths may differ (one source table each)."""
        return DeviceTableGroup(parts)


class DeviceTableGroup:
    """Several `DeviceTable`s behind one global row index (the pretraining set of the reference is a ConcatDataset over the
    source tables, scripts/train_model_with_multimodal.py:640-650).  A batch is one preprocessing launch per table that
    contributes rows, then one gather that restores the requested order."""

    def __init__(self, parts):
        self.parts = list(parts)
        assert self.parts, "at least one table"
        labelled = [p.labels is not None for p in self.parts]
        assert all(labelled) or not any(labelled), "either every table carries labels or none does"
        self.labelled = all(labelled)
        self.starts = np.cumsum([0] + [len(p) for p in self.parts])
        self.device = self.parts[0].device

    def __len__(self):
        return int(self.starts[-1])

    def batch(self, indices):
        idx = np.asarray(list(indices), dtype=np.int64)
        which = np.searchsorted(self.starts, idx, side="right") - 1
        outs, pos = [], []
        for t, part in enumerate(self.parts):
            sel = np.nonzero(which == t)[0]
            if sel.size:
                outs.append(part.batch((idx[sel] - self.starts[t]).tolist()))
                pos.append(sel)
        if not outs:
            return self.parts[0].batch([])
        back = torch.as_tensor(np.argsort(np.concatenate(pos), kind="stable")).to(self.device)
        return tuple(torch.cat([o[k] for o in outs])[back] for k in range(len(outs[0])))
