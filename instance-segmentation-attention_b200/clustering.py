"""GPU embedding clustering behind the C-ABI: the drop-in for the scikit-learn call of
`Prediction.cluster` (/root/reference/code/lib/prediction.py:52-85).

    labels = KMeans(n_clusters=n_objects, n_init=35, max_iter=500).fit_predict(X)

becomes one k-means++ seeding kernel + one persistent Lloyd kernel for all 35 restarts, with
the same seeds stream (numpy RandomState(seed) draws, made on the host: a few hundred doubles)
and the same iteration rule; see csrc/kmeans.cu for the arithmetic contract.  The reference
never seeds KMeans (random_state=None); here `seed` is explicit so runs are reproducible.
"""
import math

import numpy as np
import torch

from . import _lib


def n_local_trials(k):
    # sklearn/cluster/_kmeans.py:226: 2 + int(np.log(n_clusters))
    return 2 + int(math.log(k))


def draw_uniforms(seed, n_init, k):
    """The doubles KMeans consumes from RandomState(seed): per restart 1 (RandomState.choice for
    the first centre) + (k-1)*n_local_trials (RandomState.uniform), in stream order."""
    per = 1 + (k - 1) * n_local_trials(k)
    return np.random.RandomState(seed).random_sample(n_init * per).reshape(n_init, per)


_UNIFORMS_ON_DEVICE = {}


def _device_uniforms(seed, n_init, k, dev):
    """The same stream for the same (seed, n_init, k): uploaded once per device and reused (17 KB)."""
    key = (int(seed), int(n_init), int(k), str(dev))
    u = _UNIFORMS_ON_DEVICE.get(key)
    if u is None:
        host = torch.from_numpy(np.ascontiguousarray(draw_uniforms(seed, n_init, k), dtype=np.float64))
        u = host.to(dev)
        if len(_UNIFORMS_ON_DEVICE) > 64:
            _UNIFORMS_ON_DEVICE.clear()
        _UNIFORMS_ON_DEVICE[key] = u
    return u


class KMeansResult(object):
    __slots__ = ("labels", "centers", "inertia", "n_iter", "seed_idx", "info", "n_max")

    def check(self):
        """Synchronises and raises like scikit-learn does for bad input."""
        info = self.info.cpu().numpy()
        if info[0] == 1:
            raise ValueError("n_samples=%d should be >= n_clusters." % int(info[2]))
        if info[0] == 2:
            raise ValueError("Input X contains NaN or infinity.")
        return self

    @property
    def best(self):
        return int(self.info[1])

    @property
    def n(self):
        return int(self.info[2])


def kmeans_fit(Xt, n_dev, k, seed=0, n_init=35, max_iter=500, tol=1e-4, init_centers=None, uniforms=None):
    """Xt: (C, ld) float32 CUDA, feature-major; the first n columns are the points, n = n_dev[0]
    (device int32 tensor, so the count produced by the fg compaction never visits the host).
    Returns a KMeansResult of device tensors (labels int32 (ld,), valid up to n)."""
    lib = _lib.load()
    _lib.require_cuda(Xt, "Xt")
    assert Xt.dtype == torch.float32 and Xt.dim() == 2 and Xt.is_contiguous()
    C, ld = Xt.shape
    dev = Xt.device
    k = int(k)
    L = n_local_trials(k)
    u_dev = None
    ic = None
    if init_centers is not None:
        ic = torch.as_tensor(init_centers, dtype=torch.float32).to(dev).contiguous()
        assert tuple(ic.shape) == (n_init, k, C)
    else:
        if uniforms is None:
            u_dev = _device_uniforms(seed, n_init, k, dev)
        else:
            u_dev = torch.from_numpy(np.ascontiguousarray(uniforms, dtype=np.float64)).to(dev)
    res = KMeansResult()
    res.labels = torch.empty(ld, device=dev, dtype=torch.int32)
    res.centers = torch.empty(k, C, device=dev, dtype=torch.float32)
    res.inertia = torch.empty(n_init, device=dev, dtype=torch.float64)
    res.n_iter = torch.empty(n_init, device=dev, dtype=torch.int32)
    res.seed_idx = torch.empty(n_init, k, device=dev, dtype=torch.int32)
    res.info = torch.zeros(16, device=dev, dtype=torch.int32)
    res.n_max = ld
    wsb = lib.isa_kmeans_workspace_bytes(ld, C, k, n_init)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    rc = lib.isa_kmeans_fit(_lib.ptr(Xt), _lib.ptr(n_dev), ld, C, k, n_init, max_iter, float(tol), L,
                            _lib.ptr(u_dev), _lib.ptr(ic), _lib.ptr(res.labels), _lib.ptr(res.centers),
                            _lib.ptr(res.inertia), _lib.ptr(res.n_iter), _lib.ptr(res.seed_idx), _lib.ptr(res.info),
                            _lib.ptr(ws), wsb, _lib.stream_ptr(dev))
    _lib.check(rc, "isa_kmeans_fit")
    return res


def kmeans_fit_predict(X, k, seed=0, n_init=35, max_iter=500, tol=1e-4, init_centers=None):
    """scikit-learn-shaped convenience: X (n, C) float32 CUDA tensor -> labels (n,) int32 (device)."""
    _lib.require_cuda(X, "X")
    n, C = X.shape
    Xt = X.t().contiguous()
    n_dev = torch.tensor([n], device=X.device, dtype=torch.int32)
    res = kmeans_fit(Xt, n_dev, k, seed, n_init, max_iter, tol, init_centers).check()
    return res.labels[:n], res


def fg_compact(sem, emb):
    """sem (ncls,h,w), emb (C,h,w) float32 CUDA -> (cls_map u8 (h,w), Xt (C,h*w), fg_index i32 (h*w,), n (1,) i32)."""
    lib = _lib.load()
    _lib.require_cuda(sem, "sem")
    _lib.require_cuda(emb, "emb")
    sem = sem.contiguous().float()
    emb = emb.contiguous().float()
    ncls, h, w = sem.shape
    C = emb.shape[0]
    HW = h * w
    dev = emb.device
    cls_map = torch.empty(h, w, device=dev, dtype=torch.uint8)
    Xt = torch.empty(C, HW, device=dev, dtype=torch.float32)
    fg_index = torch.empty(HW, device=dev, dtype=torch.int32)
    n_dev = torch.empty(1, device=dev, dtype=torch.int32)
    wsb = lib.isa_fg_compact_workspace_bytes(HW)
    ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
    rc = lib.isa_fg_compact(_lib.ptr(sem), _lib.ptr(emb), ncls, C, HW, HW, _lib.ptr(cls_map), _lib.ptr(Xt),
                            _lib.ptr(fg_index), _lib.ptr(n_dev), _lib.ptr(ws), wsb, _lib.stream_ptr(dev))
    _lib.check(rc, "isa_fg_compact")
    return cls_map, Xt, fg_index, n_dev


def scatter_labels_upsample(labels, fg_index, n_dev, cls_map, out_h=None, out_w=None):
    """-> (ins_small u8 (h,w), ins_up u8 (out_h,out_w) or None, cls_up u8 or None)."""
    lib = _lib.load()
    h, w = cls_map.shape
    dev = cls_map.device
    ins_small = torch.empty(h, w, device=dev, dtype=torch.uint8)
    ins_up = cls_up = None
    if out_h is not None:
        ins_up = torch.empty(out_h, out_w, device=dev, dtype=torch.uint8)
        cls_up = torch.empty(out_h, out_w, device=dev, dtype=torch.uint8)
    rc = lib.isa_scatter_labels_upsample(_lib.ptr(labels), _lib.ptr(fg_index), _lib.ptr(n_dev), _lib.ptr(cls_map),
                                         h, w, out_h or h, out_w or w, _lib.ptr(ins_small), _lib.ptr(ins_up),
                                         _lib.ptr(cls_up), _lib.stream_ptr(dev))
    _lib.check(rc, "isa_scatter_labels_upsample")
    return ins_small, ins_up, cls_up


def cluster_embeddings(sem, emb, n_objects, out_h=None, out_w=None, seed=0, n_init=35, max_iter=500):
    """Device-resident Prediction.cluster (+ upsample_prediction): returns
    (cls_map, ins_small, ins_up, cls_up, KMeansResult) -- all CUDA tensors, nothing synchronised."""
    cls_map, Xt, fg_index, n_dev = fg_compact(sem, emb)
    res = kmeans_fit(Xt, n_dev, int(n_objects), seed, n_init, max_iter)
    ins_small, ins_up, cls_up = scatter_labels_upsample(res.labels, fg_index, n_dev, cls_map, out_h, out_w)
    return cls_map, ins_small, ins_up, cls_up, res
