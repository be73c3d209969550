"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (imported from
/root/reference through oracle/ref_loader.py) and scikit-learn / OpenCV (the reference's
third-party arithmetic) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py [disc|attn|kmeans|resize|all]

The GPU box has no /root/reference: tests read only the committed .npz files.
Inputs are regenerated from seeds by isa_b200.synth (not stored) unless tiny.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from isa_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

DISC_CASES = [
    # name, seed, bs, C, H, W, K, norm, kwargs for synth.batch
    ("a", 0, 2, 24, 48, 40, 32, 2, dict()),
    ("b", 1, 3, 8, 33, 37, 4, 2, dict(n_max=4)),
    ("c", 2, 2, 32, 40, 64, 16, 1, dict(n_max=16)),
    ("d", 3, 1, 16, 64, 64, 32, 2, dict(n_min=1, n_max=1)),
]


def gen_disc():
    ref = ref_loader.discriminative()
    out = {}
    for name, seed, bs, C, H, W, K, norm, kw in DISC_CASES:
        d = synth.batch(seed, bs, C, H, W, K, **kw)
        emb = torch.tensor(d["emb"], requires_grad=True)
        tgt = torch.tensor(synth.onehot(d["labels"], K))
        nobj = torch.tensor(d["n_objects"].astype(np.int64))
        crit = ref.DiscriminativeLoss(0.5, 1.5, norm, usegpu=False)
        loss, means = crit(emb, tgt, nobj, K)
        # a downstream use of the returned means, so their gradient path is pinned too
        gm = torch.tensor(np.random.RandomState(100 + seed).standard_normal(means.shape).astype(np.float32))
        total = loss + (means * gm).sum() * 0.01
        total.backward()
        out["%s_loss" % name] = loss.detach().numpy()
        out["%s_means" % name] = means.detach().numpy()
        out["%s_grad" % name] = emb.grad.numpy()
        out["%s_gm" % name] = gm.numpy() * 0.01
        # individual terms of the reference (distance / regulariser are defined but not summed)
        x = emb.detach().permute(0, 2, 3, 1).contiguous().view(bs, H * W, C)
        t = tgt.permute(0, 2, 3, 1).contiguous().view(bs, H * W, K)
        var = ref.calculate_variance_term(x, t, means.detach(), nobj, 0.5, norm)
        dist = ref.calculate_distance_term(means.detach(), nobj, 1.5, norm, usegpu=False)
        reg = ref.calculate_regularization_term(means.detach(), nobj, norm)
        qreg = ref.calculate_q_regularization_term(x, t)
        out["%s_terms" % name] = np.array([float(var), float(dist), float(reg), float(qreg)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "disc_loss.npz"), **out)
    print("disc_loss.npz:", {k: v.shape for k, v in out.items() if k.endswith("loss") or k.endswith("terms")})


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    assert ref_loader.available(), "reference tree missing"
    torch.manual_seed(0)
    torch.set_num_threads(1)
    if what in ("disc", "all"):
        gen_disc()
    for extra in ("attn", "kmeans", "resize", "renet", "spatial"):
        if what in (extra, "all"):
            mod = os.path.join(HERE, "make_golden_%s.py" % extra)
            if os.path.exists(mod):
                import runpy
                runpy.run_path(mod, run_name="__main__")
