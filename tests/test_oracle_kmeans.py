"""CPU: the deterministic k-means oracle (oracle/kmeans_oracle.c) pinned against golden labels
from the REAL scikit-learn 1.9.0 (tests/golden/kmeans.npz), the cv2 nearest-resize rule, and the
evaluate.py metric restatement."""
import os
import sys

import numpy as np
import pytest

from oracle import kmeans as KM

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_kmeans import KM_CASES, NET_CASES, case_inputs, net_inputs  # noqa: E402


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "kmeans.npz"))


@pytest.mark.parametrize("case", KM_CASES, ids=[c[0] for c in KM_CASES])
def test_lloyd_from_sklearn_seeds_matches_sklearn(golden, case):
    """Same seeds (scikit-learn's own k-means++ centres) -> identical partition."""
    name, seed, C, H, W, k, pull, n_init, km_seed = case
    sem, emb = case_inputs(case)
    _, X = KM.gather_foreground(sem, emb)
    inits = golden[name + "_inits"]
    for r in range(inits.shape[0]):
        o = KM.kmeans_oracle(X, k, n_init=1, init_centers=inits[r:r + 1])
        assert KM.same_up_to_permutation(o["labels"], golden[name + "_init_labels"][r]), (name, r)


@pytest.mark.parametrize("case", KM_CASES, ids=[c[0] for c in KM_CASES])
def test_full_fit_matches_sklearn(golden, case):
    """Whole KMeans(n_init, random_state=seed): same RandomState stream, same rule -> same partition."""
    name, seed, C, H, W, k, pull, n_init, km_seed = case
    sem, emb = case_inputs(case)
    _, X = KM.gather_foreground(sem, emb)
    o = KM.kmeans_oracle(X, k, seed=km_seed, n_init=n_init)
    assert KM.same_up_to_permutation(o["labels"], golden[name + "_sk_labels"]), name


@pytest.mark.parametrize("case", NET_CASES, ids=[c[0] for c in NET_CASES])
def test_network_embeddings_match_sklearn(golden, case):
    """Near-tied embeddings of a random-init network (dumped on the B200): the picks of k-means++ depend on the float32
    rounding of scikit-learn's sequential cumulative sums there.  The oracle draws scikit-learn's seeds and ends in
    scikit-learn's partition."""
    name, k, n_init, km_seed = case
    X = net_inputs(name)
    o = KM.kmeans_oracle(X, k, seed=km_seed, n_init=n_init)
    assert np.array_equal(o["seed_idx"][0], golden[name + "_sk_first_seeds"]), "k-means++ picks of the first restart"
    assert KM.same_up_to_permutation(o["labels"], golden[name + "_sk_labels"]), name


def test_seeding_is_numpy_restatement_of_sklearn():
    """isa_km_oracle_seed == _kmeans_plusplus restated with scikit-learn's own distance function and numpy's float32
    cumsum, with the potentials rounded from exact sums (the one quantity scikit-learn leaves to BLAS)."""
    from sklearn.metrics.pairwise import _euclidean_distances
    X = net_inputs("net0")[:3000]
    k, R = 8, 4
    n = len(X)
    Xc = X - X.astype(np.float64).mean(axis=0).astype(np.float32)
    o = KM.kmeans_oracle(X, k, seed=5, n_init=R, max_iter=1)
    rs = np.random.RandomState(5)
    L = KM.n_local_trials(k)
    for r in range(R):
        # RandomState.choice(n, p=uniform) consumes one double: searchsorted(cumsum(p), u, side='right')
        u = rs.random_sample()
        c0 = int(np.searchsorted(np.arange(1, n + 1) / n, u, side="right"))
        picks = [min(c0, n - 1)]
        cd = _euclidean_distances(Xc[picks[0]][None], Xc, squared=True)
        pot = np.float32(cd.astype(np.float64).sum())
        for c in range(1, k):
            rv = rs.random_sample(L) * pot
            cand = np.minimum(np.searchsorted(np.cumsum(cd[0]), rv), n - 1)
            d = np.minimum(cd, _euclidean_distances(Xc[cand], Xc, squared=True))
            cp = d.astype(np.float64).sum(axis=1).astype(np.float32)
            b = int(np.argmin(cp))
            pot, cd = cp[b], d[b][None]
            picks.append(int(cand[b]))
        assert picks == list(o["seed_idx"][r]), r


def test_uniform_stream_is_what_sklearn_consumes():
    """kmeans++ picks of the oracle == scikit-learn's for the first restart on an easy case
    (validates the choice()/uniform() stream bookkeeping)."""
    from sklearn.cluster import kmeans_plusplus
    rs = np.random.RandomState(0)
    X = rs.standard_normal((500, 4)).astype(np.float32) + 10 * rs.randint(0, 4, size=(500, 1))
    Xc = X - X.mean(0)
    _, idx = kmeans_plusplus(Xc, 4, random_state=7)
    o = KM.kmeans_oracle(X, 4, seed=7, n_init=1)
    assert np.array_equal(idx, o["seed_idx"][0])


def test_errors_and_edges():
    X = np.zeros((3, 2), dtype=np.float32)
    with pytest.raises(ValueError):
        KM.kmeans_oracle(X, 5)
    bad = np.array([[np.nan, 0], [1, 1], [2, 2]], dtype=np.float32)
    with pytest.raises(ValueError):
        KM.kmeans_oracle(bad, 2)
    # duplicates: more clusters than distinct points exercises the empty-cluster path
    X = np.repeat(np.array([[0, 0], [1, 1], [5, 5]], dtype=np.float32), 20, axis=0)
    o = KM.kmeans_oracle(X, 5, seed=0, n_init=3)
    assert o["labels"].shape == (60,) and len(np.unique(o["labels"])) <= 5
    # k == 1
    o = KM.kmeans_oracle(np.random.RandomState(0).standard_normal((50, 3)).astype(np.float32), 1, n_init=2)
    assert np.all(o["labels"] == 0)


def test_nearest_rule_matches_cv2():
    import cv2
    rs = np.random.RandomState(0)
    for src, dst in [(256, 530), (256, 500), (64, 1000), (1024, 2048)] + [(int(rs.randint(1, 300)), int(rs.randint(1, 900))) for _ in range(40)]:
        img = np.arange(src, dtype=np.float32).reshape(1, src).repeat(2, 0)
        out = cv2.resize(img, (dst, 2), interpolation=cv2.INTER_NEAREST)[0].astype(np.int64)
        assert np.array_equal(out, KM.nearest_index_map(src, dst)), (src, dst)


def test_cluster_reference_roundtrip():
    case = KM_CASES[1]
    sem, emb = case_inputs(case)
    fg, mask_sk = KM.cluster_reference(sem, emb, case[5], seed=case[8], n_init=5, impl="sklearn")
    _, mask_or = KM.cluster_reference(sem, emb, case[5], seed=case[8], n_init=5, impl="oracle")
    assert mask_sk.dtype == np.uint8 and mask_sk.shape == fg.shape
    assert np.array_equal(mask_sk == 0, fg == 0)
    assert KM.same_up_to_permutation(mask_sk, mask_or)
