"""KNN evaluation of embeddings on the GPU: the scikit-learn calls of the reference's stage-3 evaluation
(scripts/train_model_with_multimodal.py:916-934) behind the same names.

    knn = KNeighborsClassifier(n_neighbors=k).fit(emb_train, y_train); pred = knn.predict(emb_test)
    balanced_accuracy_score(y_val, pred); confusion_matrix(y_val, pred)
    knn_sweep(emb_train, y_train, emb_test, y_val, range(5, 20))      # the whole loop in three kernel launches

Euclidean metric and uniform weights only (what the reference uses); k <= 32.  Everything runs through
hippie_knn_neighbors / hippie_knn_evaluate of libhippie_b200.so on the current CUDA device; there is no CPU path --
without a GPU these functions raise.  Labels may be any sortable values: they are mapped to dense class indices over
the sorted union of the training and the true labels (np.unique, i.e. what LabelEncoder does) on the host, which is the
only host work besides moving the arrays.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib

MAX_K = 32


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("hippie_b200.knn needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _f32(x, dev) -> torch.Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32))
    t = t.to(device=dev, dtype=torch.float32).contiguous()
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D array of embeddings, got shape {tuple(t.shape)}")
    return t


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def kneighbors(train, query, k: int, return_distance: bool = True):
    """Device tensors (sqdist float64 [n_query, k] or None, index int64 [n_query, k])."""
    dev = _device()
    t, q = _f32(train, dev), _f32(query, dev)
    if t.shape[1] != q.shape[1]:
        raise ValueError(f"train has {t.shape[1]} features, query {q.shape[1]}")
    if not 1 <= k <= MAX_K:
        raise ValueError(f"n_neighbors must be in [1, {MAX_K}], got {k}")
    if k > t.shape[0]:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {k}, n_samples_fit = {t.shape[0]}")
    idx = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
    d2 = torch.empty((q.shape[0], k), dtype=torch.float64, device=dev) if return_distance else None
    rc = _lib.lib().hippie_knn_neighbors(_ptr(t), t.shape[0], _ptr(q), q.shape[0], t.shape[1], k, _ptr(idx), _ptr(d2),
                                         _stream())
    if rc:
        raise RuntimeError(f"hippie_knn_neighbors failed ({rc})")
    return d2, idx


def _evaluate(idx: torch.Tensor, train_class: torch.Tensor, true_class: Optional[torch.Tensor], n_classes: int,
              k_lo: int, k_hi: int):
    dev = idx.device
    nk, nq = k_hi - k_lo + 1, idx.shape[0]
    pred = torch.empty((nk, nq), dtype=torch.int64, device=dev)
    cm = torch.empty((nk, n_classes, n_classes), dtype=torch.int64, device=dev) if true_class is not None else None
    acc = torch.empty((nk,), dtype=torch.float64, device=dev) if true_class is not None else None
    rc = _lib.lib().hippie_knn_evaluate(_ptr(idx), nq, idx.shape[1], _ptr(train_class), _ptr(true_class), n_classes,
                                        k_lo, k_hi, _ptr(pred), _ptr(cm), _ptr(acc), _stream())
    if rc:
        raise RuntimeError(f"hippie_knn_evaluate failed ({rc})")
    return pred, cm, acc


class KNeighborsClassifier:
    """sklearn.neighbors.KNeighborsClassifier(n_neighbors) as the reference uses it: fit / predict / kneighbors /
    score, Euclidean, uniform weights.  `classes_` as in sklearn."""

    def __init__(self, n_neighbors: int = 5):
        self.n_neighbors = int(n_neighbors)

    def fit(self, X, y):
        dev = _device()
        self._fit_X = _f32(X, dev)
        y = np.asarray(y.cpu() if isinstance(y, torch.Tensor) else y)
        if y.shape[0] != self._fit_X.shape[0]:
            raise ValueError(f"X has {self._fit_X.shape[0]} rows, y {y.shape[0]}")
        self.classes_, dense = np.unique(y, return_inverse=True)
        if len(self.classes_) > 128:
            raise ValueError("at most 128 classes")
        self._y = torch.from_numpy(dense.astype(np.int64).reshape(-1)).to(dev)
        return self

    def kneighbors(self, X, n_neighbors: Optional[int] = None, return_distance: bool = True):
        d2, idx = kneighbors(self._fit_X, X, n_neighbors or self.n_neighbors, return_distance)
        if return_distance:
            return d2.sqrt().cpu().numpy(), idx.cpu().numpy()
        return idx.cpu().numpy()

    def predict(self, X):
        _, idx = kneighbors(self._fit_X, X, self.n_neighbors, return_distance=False)
        pred, _, _ = _evaluate(idx, self._y, None, len(self.classes_), self.n_neighbors, self.n_neighbors)
        return self.classes_[pred[0].cpu().numpy()]

    def score(self, X, y):
        return float(np.mean(self.predict(X) == np.asarray(y)))


def _dense(y_true, y_pred):
    labels = np.unique(np.concatenate([np.asarray(y_true).reshape(-1), np.asarray(y_pred).reshape(-1)]))
    return labels, np.searchsorted(labels, np.asarray(y_true)), np.searchsorted(labels, np.asarray(y_pred))


def confusion_matrix(y_true, y_pred) -> np.ndarray:
    """sklearn.metrics.confusion_matrix(y_true, y_pred): rows = true, columns = predicted, over the sorted union of
    the labels.  Counted on the device (a one-neighbour vote over an identity table reuses the evaluation kernel)."""
    dev = _device()
    labels, t, p = _dense(y_true, y_pred)
    if len(labels) > 128:
        raise ValueError("at most 128 classes")
    n = len(t)
    idx = torch.arange(n, dtype=torch.int64, device=dev).reshape(n, 1)
    _, cm, _ = _evaluate(idx, torch.from_numpy(p.astype(np.int64)).to(dev), torch.from_numpy(t.astype(np.int64)).to(dev),
                         len(labels), 1, 1)
    return cm[0].cpu().numpy()


def balanced_accuracy_score(y_true, y_pred) -> float:
    """sklearn.metrics.balanced_accuracy_score(y_true, y_pred): mean recall over the classes present in y_true."""
    dev = _device()
    labels, t, p = _dense(y_true, y_pred)
    if len(labels) > 128:
        raise ValueError("at most 128 classes")
    n = len(t)
    idx = torch.arange(n, dtype=torch.int64, device=dev).reshape(n, 1)
    _, _, acc = _evaluate(idx, torch.from_numpy(p.astype(np.int64)).to(dev), torch.from_numpy(t.astype(np.int64)).to(dev),
                          len(labels), 1, 1)
    return float(acc[0].item())


def knn_sweep(emb_train, y_train, emb_test, y_test, neighbor_options: Iterable[int] = range(5, 20)):
    """The reference's evaluation loop (scripts/train_model_with_multimodal.py:916-934) in one pass: neighbours once for
    max(k), then votes, confusion matrices and balanced accuracies for every k.  Returns a dict with `neighbor_options`,
    `balanced_accuracy` (list of float, one per k), `best_neighbors` (first arg-max, as np.argmax), `pred` (labels at the
    best k), `pred_all` [nk, n_test], `confusion` (at the best k, over `confusion_labels` as sklearn spans it), `confusion_all`
    [nk, C, C] over `labels`."""
    ks = list(neighbor_options)
    if not ks or ks != list(range(ks[0], ks[-1] + 1)):
        raise ValueError("neighbor_options must be a contiguous ascending range")
    dev = _device()
    y_train = np.asarray(y_train.cpu() if isinstance(y_train, torch.Tensor) else y_train).reshape(-1)
    y_test = np.asarray(y_test.cpu() if isinstance(y_test, torch.Tensor) else y_test).reshape(-1)
    labels = np.unique(np.concatenate([y_train, y_test]))
    if len(labels) > 128:
        raise ValueError("at most 128 classes")
    tr = torch.from_numpy(np.searchsorted(labels, y_train).astype(np.int64)).to(dev)
    te = torch.from_numpy(np.searchsorted(labels, y_test).astype(np.int64)).to(dev)
    _, idx = kneighbors(emb_train, emb_test, ks[-1], return_distance=False)
    pred, cm, acc = _evaluate(idx, tr, te, len(labels), ks[0], ks[-1])
    acc_h = acc.cpu().numpy()
    best = int(np.argmax(acc_h))
    pred_h = pred.cpu().numpy()
    cm_h = cm.cpu().numpy()
    # sklearn's confusion_matrix(y_test, pred) spans the labels that occur in y_test or in pred, not all training classes
    seen = np.zeros(len(labels), dtype=bool)
    seen[np.searchsorted(labels, y_test)] = True
    seen[pred_h[best]] = True
    return {"neighbor_options": ks, "balanced_accuracy": [float(a) for a in acc_h], "best_neighbors": ks[best],
            "pred": labels[pred_h[best]], "pred_all": labels[pred_h], "confusion": cm_h[best][np.ix_(seen, seen)],
            "confusion_labels": labels[seen], "confusion_all": cm_h, "labels": labels, "neighbors": idx}
