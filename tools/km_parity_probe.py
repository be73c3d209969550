"""Per-restart comparison of oracle/kmeans_oracle.c with the real scikit-learn on NETWORK embeddings
(seed-23 random-init net on synth.leaf_image), run on the CPU.  Diagnostic tool, not product."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from isa_b200 import synth
from oracle import kmeans as KM
from oracle.model_ref import ReSegRef
from PIL import Image
from isa_b200.settings import CVPPPModelSettings

def embedding(idx):
    cache = "/tmp/km_probe_X%d.npy" % idx
    if os.path.exists(cache):
        return np.load(cache)
    ms_ = CVPPPModelSettings()
    torch.manual_seed(23)
    net = ReSegRef(2, n_embedding=24).eval()
    raw = synth.leaf_image(idx, 530, 500)
    img = Image.fromarray(raw).resize((256, 256), Image.BILINEAR)
    x = (np.asarray(img, dtype=np.float32).transpose(2, 0, 1) / 255.0 - np.asarray(ms_.MEAN, np.float32).reshape(3, 1, 1)) / np.asarray(ms_.STD, np.float32).reshape(3, 1, 1)
    with torch.no_grad():
        sem_out, emb = net(False, torch.from_numpy(x).unsqueeze(0))
        sem_p = torch.softmax(sem_out, 1)
    fg, X = KM.gather_foreground(sem_p[0].numpy(), emb[0].numpy())
    np.save(cache, X)
    return X

def sklearn_per_restart(X, k, seed, n_init=35, max_iter=500, tol=1e-4):
    """Replays KMeans.fit's loop (sklearn/cluster/_kmeans.py:1487-1545) restart by restart."""
    from sklearn.cluster import _kmeans as SK
    from sklearn.utils import check_random_state
    from sklearn.utils.extmath import row_norms
    from sklearn.cluster._kmeans import _kmeans_single_lloyd, _tolerance
    km = SK.KMeans(n_clusters=k, n_init=n_init, max_iter=max_iter, random_state=seed)
    rs = check_random_state(seed)
    Xc = X.copy()
    mean = Xc.mean(axis=0)
    Xc -= mean
    tol_abs = _tolerance(Xc, tol)
    xsq = row_norms(Xc, squared=True)
    sw = np.ones(len(Xc), dtype=Xc.dtype)
    km._n_threads = 1 if os.environ.get("SK1") else SK._openmp_effective_n_threads()
    km._algorithm = "lloyd"
    out = []
    for r in range(n_init):
        ci = km._init_centroids(Xc, x_squared_norms=xsq, init="k-means++", random_state=rs, sample_weight=sw)
        labels, inertia, centers, n_iter = _kmeans_single_lloyd(Xc, sw, ci, max_iter=max_iter, verbose=False, tol=tol_abs, n_threads=km._n_threads)
        out.append((labels.copy(), float(inertia), int(n_iter), ci.copy()))
    return out, tol_abs

if __name__ == "__main__":
    k = 16
    for idx in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]:
        X = embedding(idx)
        t0 = time.time()
        o = KM.kmeans_oracle(X, k, seed=0)
        t1 = time.time()
        sk, tol_abs = sklearn_per_restart(X, k, 0)
        t2 = time.time()
        full = KM.sklearn_fit_predict(X, k, 0)
        print("image %d n=%d oracle %.1fs sklearn %.1fs tol_abs sk=%.6g oracle=%.6g" % (idx, len(X), t1 - t0, t2 - t1, tol_abs, o["tol_abs"]))
        sk_best = int(np.argmin([s[1] for s in sk]))
        # sklearn picks first strictly-lower inertia
        bi, best = None, None
        for r, s in enumerate(sk):
            if best is None or s[1] < best:
                best, bi = s[1], r
        print("  oracle best=%d (n_iter %d, inertia %.6f) sklearn best=%d (n_iter %d inertia %.6f) same-as-full=%s" % (
            o["best"], o["n_iter"][o["best"]], o["inertia"][o["best"]], bi, sk[bi][2], sk[bi][1], KM.same_up_to_permutation(full, sk[bi][0])))
        print("  oracle-vs-sklearn final: identical=%s agreement=%.5f" % (KM.same_up_to_permutation(o["labels"], full), KM.partition_agreement(o["labels"], full)))
        nd = 0
        for r in range(35):
            flag = "" if o["n_iter"][r] == sk[r][2] else "  <-- n_iter differs"
            if flag or r in (o["best"], bi):
                print("   r=%2d oracle n_iter=%3d strict=%d inertia=%.6f | sklearn n_iter=%3d inertia=%.6f%s" % (r, o["n_iter"][r], o["strict"][r], o["inertia"][r], sk[r][2], sk[r][1], flag))
            nd += bool(flag)
        print("  restarts with differing n_iter: %d / 35" % nd)
        order_o = np.argsort(o["inertia"])[:4]; order_s = np.argsort([s[1] for s in sk])[:4]
        print("  lowest inertia oracle:", [(int(r), round(float(o["inertia"][r]), 5)) for r in order_o])
        print("  lowest inertia sklearn:", [(int(r), round(sk[r][1], 5)) for r in order_s])
