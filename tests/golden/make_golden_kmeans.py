"""Golden vectors of the clustering step from the REAL scikit-learn (the reference's third-party
arithmetic, prediction.py:72-74) and OpenCV.  Run through make_golden.py kmeans."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from isa_b200 import synth  # noqa: E402
from oracle import kmeans as KM  # noqa: E402

# name, seed, C, H, W, n_objects(k), pull, n_init, km_seed
KM_CASES = [
    ("leaf16", 3, 24, 96, 96, 16, 0.7, 35, 0),
    ("leaf8", 4, 8, 64, 80, 8, 0.8, 35, 1),
    ("noisy16", 5, 32, 64, 64, 16, 0.3, 10, 2),
    ("two", 6, 16, 48, 48, 2, 0.7, 35, 3),
]


# Network embeddings: foreground embeddings of the seed-23 random-init ReSeg network on synth.leaf_image(0) / (2) at the
# CVPPP size, dumped on the B200 by tools/ncu_infer.py (every 6th point, rounded to float16 to keep the fixture small;
# the float32 upcast of the stored values IS the input).  Near-tied, unlike the planted clusters above: the picks of
# k-means++ depend on the float32 rounding of scikit-learn's cumulative sums here.
NET_CASES = [("net0", 16, 35, 0), ("net2", 16, 35, 0)]


def net_inputs(name):
    d = np.load(os.path.join(HERE, "kmeans_net_inputs.npz"))
    return d[name + "_X16"].astype(np.float32)


def case_inputs(case):
    name, seed, C, H, W, k, pull, n_init, km_seed = case
    d = synth.batch(seed, 1, C, H, W, max(k, 2), n_min=k, n_max=k, pull=pull)
    lab = d["labels"][0]
    sem = np.stack([(lab == 255).astype(np.float32) * 0.8 + 0.1, (lab != 255).astype(np.float32) * 0.8 + 0.1])
    return sem, d["emb"][0]


def main():
    from sklearn.cluster import KMeans, kmeans_plusplus
    out = {}
    for case in KM_CASES:
        name, seed, C, H, W, k, pull, n_init, km_seed = case
        sem, emb = case_inputs(case)
        fg, X = KM.gather_foreground(sem, emb)
        out[name + "_sk_labels"] = KM.sklearn_fit_predict(X, k, km_seed, n_init).astype(np.uint8)
        # Lloyd pinned separately: 3 explicit seedings (from scikit-learn's own k-means++), n_init=1 each
        inits, labs = [], []
        for r in range(3):
            c, _ = kmeans_plusplus(X, k, random_state=100 + r)
            inits.append(c)
            labs.append(KMeans(n_clusters=k, init=c, n_init=1, max_iter=500).fit_predict(X))
        out[name + "_inits"] = np.stack(inits).astype(np.float32)
        out[name + "_init_labels"] = np.stack(labs).astype(np.uint8)
        print(name, X.shape, "sklearn clusters", len(np.unique(out[name + "_sk_labels"])))
    for name, k, n_init, km_seed in NET_CASES:
        X = net_inputs(name)
        out[name + "_sk_labels"] = KM.sklearn_fit_predict(X, k, km_seed, n_init).astype(np.uint8)
        _, idx = kmeans_plusplus(X - X.mean(axis=0), k, random_state=km_seed)    # the first restart's picks
        out[name + "_sk_first_seeds"] = idx.astype(np.int32)
        print(name, X.shape, "sklearn clusters", len(np.unique(out[name + "_sk_labels"])))
    np.savez_compressed(os.path.join(HERE, "kmeans.npz"), **out)


if __name__ == "__main__":
    main()
