"""TEST INFRASTRUCTURE ONLY -- Python face of the C k-means oracle (oracle/kmeans_oracle.c)
and the reference's own clustering call restated around the REAL scikit-learn.

  kmeans_oracle(X, k, seed, ...)   deterministic restatement, the thing the CUDA path must
                                   match bit for bit
  sklearn_fit_predict(X, k, seed)  what /root/reference/code/lib/prediction.py:72-74 calls
                                   (minus the removed n_jobs argument), scikit-learn 1.9.0
  cluster_reference(sem, emb, k)   Prediction.cluster of the reference
                                   (/root/reference/code/lib/prediction.py:52-85) on numpy arrays
  draw_uniforms(seed, n_init, k)   the RandomState stream KMeans consumes: per restart one
                                   double for RandomState.choice and (k-1)*(2+floor(ln k)) for
                                   RandomState.uniform (sklearn/cluster/_kmeans.py:232,246)
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libkmeans_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _LIB = ctypes.CDLL(path)
        _LIB.isa_km_oracle_fit.restype = ctypes.c_int
    return _LIB


def n_local_trials(k):
    return 2 + int(math.log(k))


def draw_uniforms(seed, n_init, k):
    per = 1 + (k - 1) * n_local_trials(k)
    return np.random.RandomState(seed).random_sample(n_init * per).reshape(n_init, per)


def kmeans_oracle(X, k, seed=0, n_init=35, max_iter=500, tol=1e-4, init_centers=None, uniforms=None):
    """X (n,C) float32.  Returns dict(labels, centers, inertia[r], n_iter[r], strict[r], seed_idx, best, tol_abs)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, C = X.shape
    if uniforms is None:
        uniforms = draw_uniforms(seed, n_init, k)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    labels = np.empty(n, dtype=np.int32)
    centers = np.empty((k, C), dtype=np.float32)
    inertia = np.empty(n_init, dtype=np.float64)
    n_iter = np.empty(n_init, dtype=np.int32)
    strict = np.empty(n_init, dtype=np.int32)
    seed_idx = np.empty((n_init, k), dtype=np.int32)
    best = ctypes.c_int(0)
    tol_abs = ctypes.c_float(0)
    ic = None
    if init_centers is not None:
        ic = np.ascontiguousarray(init_centers, dtype=np.float32)
        assert ic.shape == (n_init, k, C)
    P = ctypes.c_void_p
    st = _lib().isa_km_oracle_fit(
        X.ctypes.data_as(P), n, C, k, n_init, max_iter, ctypes.c_double(tol), uniforms.ctypes.data_as(P),
        ic.ctypes.data_as(P) if ic is not None else None,
        labels.ctypes.data_as(P), centers.ctypes.data_as(P), inertia.ctypes.data_as(P), n_iter.ctypes.data_as(P),
        strict.ctypes.data_as(P), seed_idx.ctypes.data_as(P), ctypes.byref(best), ctypes.byref(tol_abs))
    if st == 1:
        raise ValueError("n_samples=%d should be >= n_clusters=%d." % (n, k))
    if st == 2:
        raise ValueError("Input X contains NaN or infinity.")
    return dict(labels=labels, centers=centers, inertia=inertia, n_iter=n_iter, strict=strict, seed_idx=seed_idx,
                best=best.value, tol_abs=tol_abs.value)


def sklearn_fit_predict(X, k, seed=0, n_init=35, max_iter=500, init="k-means++"):
    """The reference's clustering call on the real scikit-learn (prediction.py:72-74)."""
    from sklearn.cluster import KMeans
    return KMeans(n_clusters=int(k), n_init=n_init, max_iter=max_iter, random_state=seed, init=init).fit_predict(X)


def gather_foreground(sem, emb):
    """prediction.py:57-69: fg = argmax_c(sem) != 0; X = emb[:, fg].T in np.where (row-major) order."""
    fg = sem.argmax(0).astype(np.uint8)
    e = emb.transpose(1, 2, 0)
    X = np.stack([e[:, :, i][fg != 0] for i in range(e.shape[2])], axis=1)
    return fg, np.ascontiguousarray(X, dtype=np.float32)


def scatter_labels(fg, labels):
    """prediction.py:76-83 without the python loop: mask[y,x] = label+1 at foreground pixels."""
    mask = np.zeros(fg.shape, dtype=np.uint8)
    mask[fg != 0] = (labels + 1).astype(np.uint8)
    return mask


def cluster_reference(sem, emb, k, seed=0, n_init=35, max_iter=500, impl="sklearn"):
    """Prediction.cluster (prediction.py:52-85): returns (fg uint8 (h,w), instance mask uint8 (h,w))."""
    fg, X = gather_foreground(np.asarray(sem), np.asarray(emb))
    if impl == "sklearn":
        labels = sklearn_fit_predict(X, k, seed, n_init, max_iter)
    else:
        labels = kmeans_oracle(X, k, seed, n_init, max_iter)["labels"]
    return fg, scatter_labels(fg, labels)


def upsample_nearest(mask, out_h, out_w):
    """prediction.py:47-50: cv2.resize(..., interpolation=cv2.INTER_NEAREST)."""
    import cv2
    return cv2.resize(mask, (out_w, out_h), interpolation=cv2.INTER_NEAREST)


def nearest_index_map(src, dst):
    """Integer restatement of OpenCV's INTER_NEAREST source index: min(floor(x * src/dst), src-1)
    with the scale computed in double as 1/(dst/src) (modules/imgproc/src/resize.cpp, resizeNN)."""
    fx = dst / float(src)
    ifx = 1.0 / fx
    return np.minimum(np.floor(np.arange(dst) * ifx).astype(np.int64), src - 1)


def same_up_to_permutation(a, b):
    """True when label arrays a and b describe the same partition."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    if a.shape != b.shape:
        return False
    pairs = np.unique(np.stack([a, b], 1), axis=0)
    return len(np.unique(pairs[:, 0])) == len(pairs) and len(np.unique(pairs[:, 1])) == len(pairs)


def partition_agreement(a, b):
    """Fraction of points on which label arrays a and b agree under the best one-to-one relabelling
    (1.0 == same partition).  Used to QUANTIFY a mismatch against real scikit-learn, whose own fp32 sums are
    not reproducible across BLAS builds / thread counts (near-tied restarts can then be ranked differently)."""
    from scipy.optimize import linear_sum_assignment
    a = np.asarray(a).ravel().astype(np.int64)
    b = np.asarray(b).ravel().astype(np.int64)
    ua, ia = np.unique(a, return_inverse=True)
    ub, ib = np.unique(b, return_inverse=True)
    cont = np.zeros((len(ua), len(ub)), dtype=np.int64)
    np.add.at(cont, (ia, ib), 1)
    r, c = linear_sum_assignment(-cont)
    return float(cont[r, c].sum()) / max(len(a), 1)

