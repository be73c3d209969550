"""Seeded synthetic inputs shaped like the reference's data (no dataset ships and there is
no network): Voronoi instance label maps inside a foreground ellipse, one-hot targets as the
reference collate emits them (/root/reference/code/lib/dataset.py:354-376), embeddings pulled
toward per-instance unit-norm centres so the hinges are partly active.
Used by tests/, bench.py and __graft_entry__.smoke().  numpy only, deterministic per seed.
"""
import numpy as np


def label_map(rs, H, W, n_objects, fg_frac=0.5):
    """(H,W) uint8 labels 0..n-1 inside an ellipse covering ~fg_frac of the image, 255 = background.
    Every instance id below n_objects owns at least one pixel."""
    yy, xx = np.mgrid[0:H, 0:W]
    a = np.sqrt(fg_frac * 4 / np.pi)  # ellipse semi-axes as a fraction of half-size
    fg = ((yy - (H - 1) / 2.0) / (a * H / 2.0 + 1e-9)) ** 2 + ((xx - (W - 1) / 2.0) / (a * W / 2.0 + 1e-9)) ** 2 <= 1.0
    ys, xs = np.nonzero(fg)
    if len(ys) < n_objects:
        fg[:] = True
        ys, xs = np.nonzero(fg)
    pick = rs.choice(len(ys), size=n_objects, replace=False)
    sy, sx = ys[pick].astype(np.float64), xs[pick].astype(np.float64)
    d = (yy[None] - sy[:, None, None]) ** 2 + (xx[None] - sx[:, None, None]) ** 2
    lab = d.argmin(0).astype(np.uint8)
    lab[~fg] = 255
    return lab


def batch(seed, bs, C, H, W, K, n_min=2, n_max=None, fg_frac=0.5, pull=0.7, sigma=1.0):
    """Returns dict(emb (bs,C,H,W) f32, labels (bs,H,W) u8, n_objects (bs,) i32, centres (bs,K,C))."""
    rs = np.random.RandomState(seed)
    n_max = min(K, 23) if n_max is None else min(n_max, K)
    n_min = min(n_min, n_max)
    labels = np.empty((bs, H, W), dtype=np.uint8)
    n_objects = rs.randint(n_min, n_max + 1, size=bs).astype(np.int32)
    emb = (sigma * rs.standard_normal((bs, C, H, W))).astype(np.float32)
    centres = rs.standard_normal((bs, K, C))
    centres /= np.linalg.norm(centres, axis=2, keepdims=True)
    centres = centres.astype(np.float32)
    for b in range(bs):
        labels[b] = label_map(rs, H, W, int(n_objects[b]), fg_frac)
        lab = labels[b]
        m = lab != 255
        c = centres[b][lab[m].astype(np.int64)]  # (nfg, C)
        e = emb[b][:, m]
        emb[b][:, m] = ((1.0 - pull) * e + pull * c.T).astype(np.float32)
    return dict(emb=emb, labels=labels, n_objects=n_objects, centres=centres)


def onehot(labels, K, dtype=np.float32):
    """(bs,H,W) u8 label map -> (bs,K,H,W) one-hot masks (background = all zero)."""
    bs, H, W = labels.shape
    out = np.zeros((bs, K, H, W), dtype=dtype)
    for k in range(K):
        out[:, k] = (labels == k)
    return out


def leaf_image(seed, H=530, W=500):
    """CVPPP-shaped synthetic RGB uint8 image (smoothed noise so fg/bg regions exist)."""
    rs = np.random.RandomState(seed)
    small = rs.randint(0, 256, size=(H // 16 + 2, W // 16 + 2, 3)).astype(np.float32)
    # separable linear up-sampling (cheap stand-in for a Gaussian blur, sigma ~ 8 px)
    yi = np.linspace(0, small.shape[0] - 1.001, H)
    xi = np.linspace(0, small.shape[1] - 1.001, W)
    y0, x0 = yi.astype(int), xi.astype(int)
    fy, fx = (yi - y0)[:, None, None], (xi - x0)[None, :, None]
    img = (small[y0][:, x0] * (1 - fy) * (1 - fx) + small[y0 + 1][:, x0] * fy * (1 - fx)
           + small[y0][:, x0 + 1] * (1 - fy) * fx + small[y0 + 1][:, x0 + 1] * fy * fx)
    return np.clip(img, 0, 255).astype(np.uint8)
