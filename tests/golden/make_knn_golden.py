"""Generates tests/golden/knn_*.npz with scikit-learn itself -- the library the reference calls for its stage-3
evaluation (scripts/train_model_with_multimodal.py:916-934).  Run once in the build container:

    python tests/golden/make_knn_golden.py

Cases: z = 10 (sklearn picks its KD-tree), z = 32 (brute force), a 12-class case (numpy's blocked summation in the mean
recall), and one with a class that never occurs in the true labels.  Inputs are row-z-scored float32 embeddings of
clustered points, as get_embeddings_multimodal returns them, with the class imbalance of cellexplorer-celltype."""
import os

import numpy as np
from sklearn.metrics import balanced_accuracy_score, confusion_matrix
from sklearn.neighbors import KNeighborsClassifier

HERE = os.path.dirname(os.path.abspath(__file__))


def make_case(seed, n_train, n_test, dim, class_p, drop_true=None):
    rng = np.random.default_rng(seed)
    C = len(class_p)
    centres = rng.normal(size=(C, dim)) * 0.9
    def draw(n):
        y = rng.choice(C, size=n, p=class_p)
        x = centres[y] + rng.normal(size=(n, dim))
        x = (x - x.mean(1, keepdims=True)) / x.std(1, keepdims=True)
        return x.astype(np.float32), y.astype(np.int64)
    xtr, ytr = draw(n_train)
    xte, yte = draw(n_test)
    for c in range(C):  # every class present in the training labels (LabelEncoder-dense ids)
        ytr[c] = c
    if drop_true is not None:
        yte[yte == drop_true] = (drop_true + 1) % C
    return xtr, ytr, xte, yte


def run(name, *a, **kw):
    xtr, ytr, xte, yte = make_case(*a, **kw)
    C = int(max(ytr.max(), yte.max())) + 1
    ks = list(range(5, 20))
    knn = KNeighborsClassifier(n_neighbors=19).fit(xtr, ytr)
    dist, idx = knn.kneighbors(xte)
    preds, accs, cms = [], [], []
    for k in ks:
        m = KNeighborsClassifier(n_neighbors=k).fit(xtr, ytr)
        p = m.predict(xte)
        preds.append(p)
        accs.append(balanced_accuracy_score(yte, p))
        cms.append(confusion_matrix(yte, p, labels=np.arange(C)))
    np.savez_compressed(os.path.join(HERE, name), train=xtr, train_class=ytr, query=xte, true_class=yte,
                        neighbors=idx.astype(np.int64), dist=dist, pred=np.stack(preds).astype(np.int64),
                        balanced_accuracy=np.asarray(accs, dtype=np.float64), confusion=np.stack(cms).astype(np.int64),
                        fit_method=np.array(knn._fit_method))
    print(name, knn._fit_method, "best k", ks[int(np.argmax(accs))], "acc", max(accs))


if __name__ == "__main__":
    import warnings
    warnings.simplefilter("ignore")
    p4 = [0.56, 0.29, 0.11, 0.04]  # PV / SST / Pyra / VIP shares of cellexplorer-celltype
    run("knn_z10.npz", 1, 313, 79, 10, p4)
    run("knn_z32.npz", 2, 500, 131, 32, p4)
    run("knn_c12.npz", 3, 700, 160, 10, [1 / 12] * 12)
    run("knn_absent.npz", 4, 200, 60, 10, [0.4, 0.3, 0.2, 0.1], drop_true=3)
