"""KNN evaluation (reference scripts/train_model_with_multimodal.py:916-934).

CPU part: oracle/knn_oracle.py against scikit-learn's own outputs frozen in tests/golden/knn_*.npz
(tests/golden/make_knn_golden.py).  GPU part: hippie_knn_neighbors / hippie_knn_evaluate through the C ABI against the
fixtures and the oracle, bit-exact (indices, predictions, counts, float64 balanced accuracy)."""
import os

import numpy as np
import pytest
import torch

from oracle import knn_oracle as K

GOLDEN = ["knn_z10", "knn_z32", "knn_c12", "knn_absent"]


def _load(golden_dir, tag):
    return np.load(os.path.join(golden_dir, tag + ".npz"))


@pytest.mark.parametrize("tag", GOLDEN)
def test_oracle_matches_sklearn_fixture(golden_dir, tag):
    g = _load(golden_dir, tag)
    C = g["confusion"].shape[1]
    pred, cm, acc, idx = K.evaluate(g["train"], g["train_class"], g["query"], g["true_class"], C, 5, 19)
    assert np.array_equal(idx, g["neighbors"])
    assert np.array_equal(pred, g["pred"])
    assert np.array_equal(cm, g["confusion"])
    assert np.array_equal(acc, g["balanced_accuracy"])  # float64, bit for bit
    _, d2 = K.kneighbors(g["train"], g["query"], 19)
    if str(g["fit_method"]) == "kd_tree":  # the tree evaluates exactly this sum; brute force uses the GEMM identity
        assert np.array_equal(np.sqrt(d2), g["dist"])
    else:
        assert np.allclose(np.sqrt(d2), g["dist"], rtol=0, atol=1e-6)


def test_oracle_vote_ties_go_to_the_smallest_class():
    lab = np.array([[2, 1, 1, 2], [3, 0, 3, 0], [1, 1, 1, 0]])
    assert K.vote(lab, 4).tolist() == [1, 0, 1]


def test_knn_api_raises_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hippie_b200 import knn
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        knn.knn_sweep(np.zeros((40, 4), np.float32), np.zeros(40, int), np.zeros((4, 4), np.float32), np.zeros(4, int))


# ----------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", GOLDEN)
def test_gpu_sweep_matches_sklearn_fixture(golden_dir, tag):
    from hippie_b200 import knn
    g = _load(golden_dir, tag)
    r = knn.knn_sweep(g["train"], g["train_class"], g["query"], g["true_class"], range(5, 20))
    assert np.array_equal(r["neighbors"].cpu().numpy(), g["neighbors"])
    assert np.array_equal(r["pred_all"], g["pred"])
    C = g["confusion"].shape[1]
    assert np.array_equal(r["confusion_all"][:, :C, :C], g["confusion"])
    assert np.array_equal(np.asarray(r["balanced_accuracy"]), g["balanced_accuracy"])
    best = int(np.argmax(g["balanced_accuracy"]))
    assert r["best_neighbors"] == 5 + best
    assert np.array_equal(r["pred"], g["pred"][best])
    # sklearn's confusion_matrix(y_true, pred) spans only the labels that occur
    seen = np.union1d(g["true_class"], g["pred"][best])
    assert np.array_equal(r["confusion"], g["confusion"][best][np.ix_(seen, seen)])


@pytest.mark.gpu
def test_gpu_classifier_api_matches_fixture(golden_dir):
    from hippie_b200 import knn
    g = _load(golden_dir, "knn_z10")
    m = knn.KNeighborsClassifier(n_neighbors=19).fit(g["train"], g["train_class"])
    dist, idx = m.kneighbors(g["query"])
    assert np.array_equal(idx, g["neighbors"]) and np.array_equal(dist, g["dist"])
    for i, k in enumerate(range(5, 20)):
        p = knn.KNeighborsClassifier(n_neighbors=k).fit(g["train"], g["train_class"]).predict(g["query"])
        assert np.array_equal(p, g["pred"][i])
        assert knn.balanced_accuracy_score(g["true_class"], p) == g["balanced_accuracy"][i]
        seen = np.union1d(g["true_class"], p)
        assert np.array_equal(knn.confusion_matrix(g["true_class"], p), g["confusion"][i][np.ix_(seen, seen)])
    # string labels, as LabelEncoder.inverse_transform hands them out
    names = np.array(["PV", "SST", "Pyra", "VIP"])
    p = knn.KNeighborsClassifier(7).fit(g["train"], names[g["train_class"]]).predict(g["query"])
    # dense ids follow the SORTED names, so vote ties resolve by name order: restate that with the oracle's vote
    idx7 = g["neighbors"][:, :7]
    want = np.sort(names)[K.vote(np.searchsorted(np.sort(names), names[g["train_class"]])[idx7], 4)]
    assert np.array_equal(p, want)


@pytest.mark.gpu
@pytest.mark.parametrize("n_train,n_query,dim,k", [(5000, 777, 10, 19), (3000, 300, 64, 32), (40, 9, 1, 5),
                                                   (1500, 100, 128, 19), (33, 1, 3, 33 - 1)])
def test_gpu_neighbors_match_oracle(n_train, n_query, dim, k):
    """Several shared-memory tiles per query, ragged last tile and CTA, maximum k, one feature."""
    from hippie_b200 import knn
    rng = np.random.default_rng(n_train + dim)
    train = rng.normal(size=(n_train, dim)).astype(np.float32)
    query = rng.normal(size=(n_query, dim)).astype(np.float32)
    d2, idx = knn.kneighbors(train, query, k)
    want_idx, want_d2 = K.kneighbors(train, query, k)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert np.array_equal(d2.cpu().numpy(), want_d2)


@pytest.mark.gpu
def test_gpu_ties_duplicates_and_edges():
    from hippie_b200 import knn
    rng = np.random.default_rng(0)
    base = rng.normal(size=(50, 6)).astype(np.float32)
    train = np.concatenate([base, base, base])            # every distance occurs three times
    query = base[:20] + np.float32(0.25)
    _, idx = knn.kneighbors(train, query, 19)
    want, _ = K.kneighbors(train, query, 19)              # stable sort: equal distances by ascending index
    assert np.array_equal(idx.cpu().numpy(), want)
    # the query set itself: nearest neighbour of a training row is that row, at distance 0
    d2, idx = knn.kneighbors(base, base, 1)
    assert np.array_equal(idx.cpu().numpy()[:, 0], np.arange(50)) and float(d2.abs().max()) == 0.0
    # empty query set
    _, idx = knn.kneighbors(base, np.zeros((0, 6), np.float32), 5)
    assert idx.shape == (0, 5)
    with pytest.raises(ValueError):
        knn.kneighbors(base, base, 33)
    with pytest.raises(ValueError):
        knn.kneighbors(base[:4], base, 5)
    with pytest.raises(ValueError):
        knn.kneighbors(base, base[:, :3], 5)


@pytest.mark.gpu
def test_gpu_sweep_large_random_against_oracle():
    from hippie_b200 import knn
    rng = np.random.default_rng(7)
    C = 9
    train = rng.normal(size=(20000, 10)).astype(np.float32)
    ytr = rng.integers(0, C, size=20000)
    query = rng.normal(size=(4000, 10)).astype(np.float32)
    yte = rng.integers(0, C, size=4000)
    r = knn.knn_sweep(train, ytr, query, yte, range(5, 20))
    pred, cm, acc, idx = K.evaluate(train, ytr, query, yte, C, 5, 19)
    assert np.array_equal(r["neighbors"].cpu().numpy(), idx)
    assert np.array_equal(r["pred_all"], pred) and np.array_equal(r["confusion_all"], cm)
    assert np.array_equal(np.asarray(r["balanced_accuracy"]), acc)
    assert int(r["confusion_all"].sum()) == 15 * 4000  # every query counted once per k
