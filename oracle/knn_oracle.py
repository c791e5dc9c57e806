"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's stage-3 KNN evaluation.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(hippie_b200/knn.py -> libhippie_b200.so) never does.

Path restated: scripts/train_model_with_multimodal.py:916-934 of the reference --
    KNeighborsClassifier(n_neighbors=k).fit(emb_train, y_train).predict(emb_test)    for k in range(5, 20)
    balanced_accuracy_score(y_val, pred);  confusion_matrix(y_val, pred)
The arithmetic lives in scikit-learn (unpinned in the reference's requirements.txt:8; 1.9.0 in this image): Euclidean
metric, uniform weights, KD-tree for dim <= 15 and k < n_train // 2, brute force otherwise
(sklearn/neighbors/_base.py:_fit), both on float64 copies of the float32 embeddings.  Pinned: tests/golden/knn_*.npz hold
scikit-learn's own outputs on seeded inputs (tests/golden/make_knn_golden.py); tests/test_oracle_golden.py checks this
restatement against them bit for bit (indices, predictions, confusion counts, balanced accuracy as float64).
"""
from __future__ import annotations

import numpy as np


def sq_distances(train: np.ndarray, query: np.ndarray) -> np.ndarray:
    """[n_query, n_train] squared Euclidean distances as sklearn's KD-tree evaluates them
    (sklearn/metrics/_dist_metrics.pyx: euclidean rdist -- `tmp = x1[j] - x2[j]; d += tmp * tmp` over float64 copies)."""
    t = np.asarray(train, dtype=np.float64)
    q = np.asarray(query, dtype=np.float64)
    d = np.zeros((q.shape[0], t.shape[0]), dtype=np.float64)
    for j in range(t.shape[1]):  # feature order matters for the rounding
        diff = q[:, j:j + 1] - t[None, :, j]
        d += diff * diff
    return d


def kneighbors(train, query, k: int):
    """Indices [n_query, k] and squared distances of the k nearest train rows, ascending, ties by ascending index."""
    d = sq_distances(train, query)
    order = np.argsort(d, axis=1, kind="stable")[:, :k]
    return order.astype(np.int64), np.take_along_axis(d, order, axis=1)


def vote(neigh_classes: np.ndarray, n_classes: int) -> np.ndarray:
    """Majority class per row, ties to the smallest class (sklearn/utils/extmath.py:_mode / scipy.stats.mode)."""
    counts = np.zeros((neigh_classes.shape[0], n_classes), dtype=np.int64)
    for c in range(n_classes):
        counts[:, c] = (neigh_classes == c).sum(axis=1)
    return counts.argmax(axis=1).astype(np.int64)


def confusion(true_class, pred_class, n_classes: int) -> np.ndarray:
    cm = np.zeros((n_classes, n_classes), dtype=np.int64)
    np.add.at(cm, (np.asarray(true_class), np.asarray(pred_class)), 1)
    return cm


def balanced_accuracy(cm: np.ndarray) -> float:
    """sklearn/metrics/_classification.py:balanced_accuracy_score -- mean recall over classes with support."""
    rows = cm.sum(axis=1)
    keep = rows > 0
    if not keep.any():
        return float("nan")
    per_class = np.diag(cm)[keep] / rows[keep]
    return float(np.mean(per_class))


def evaluate(train, train_class, query, true_class, n_classes: int, k_lo: int = 5, k_hi: int = 19):
    """The whole sweep: predictions [nk, n_query], confusion [nk, C, C], balanced accuracy [nk]."""
    idx, _ = kneighbors(train, query, k_hi)
    lab = np.asarray(train_class)[idx]
    preds, cms, accs = [], [], []
    for k in range(k_lo, k_hi + 1):
        p = vote(lab[:, :k], n_classes)
        cm = confusion(true_class, p, n_classes)
        preds.append(p), cms.append(cm), accs.append(balanced_accuracy(cm))
    return np.stack(preds), np.stack(cms), np.asarray(accs, dtype=np.float64), idx
