"""GPU: BASELINE.json configs[4] -- the discriminative-loss + clustering sweep (embedding width 8-32, 1-128 instances,
256^2 - 1024^2 pixels) as parity cases.  The loss is compared with an O(P C) float64 restatement of the shipped
composite (validated against oracle/disc_loss.py in tests/test_cityscapes_shape_gpu.py; the dense-mask oracle needs
P*K*C doubles, 34 GB at the top of the sweep); k-means is bit-exact against the C oracle with a bounded restart /
iteration budget.  The 1024^2 corner of the sweep has its own cases below (and 1024x2048 in tests/test_cityscapes_shape_gpu.py)."""
import numpy as np
import pytest
import torch

from isa_b200 import synth
from oracle import kmeans as KM

from .test_cityscapes_shape_gpu import _loss_ref_label_map, _rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C", [8, 16, 32])
@pytest.mark.parametrize("K,n_max", [(1, 1), (4, 4), (32, 23), (128, 128)])
@pytest.mark.parametrize("HW", [256, 512])
def test_loss_sweep(cuda, C, K, n_max, HW):
    from isa_b200.losses import DiscriminativeLoss
    if HW == 512 and C == 16:
        pytest.skip("the 512^2 column of the sweep runs at the two extreme widths")
    bs = 2 if HW == 256 else 1
    d = synth.batch(C * 7 + K, bs, C, HW, HW, K, n_min=max(n_max // 2, 1), n_max=n_max, pull=0.7)
    x = torch.tensor(d["emb"], device=cuda, requires_grad=True)
    loss, means = DiscriminativeLoss(0.5, 1.5, 2)(x, torch.tensor(d["labels"], device=cuda), torch.tensor(d["n_objects"], device=cuda), K)
    loss.backward()
    # reference: the composite is the batch mean of the per-image variance terms + 0.005 x ONE q-regulariser over the batch
    refs = [_loss_ref_label_map(d["emb"][b], d["labels"][b], int(d["n_objects"][b])) for b in range(bs)]
    if bs == 1:
        l_ref, g_ref = refs[0][0], refs[0][2][None]
        assert abs(float(loss) - l_ref) <= 1e-4 * abs(l_ref)
        assert _rel(x.grad, g_ref) < 1e-4
    for b in range(bs):
        n = int(d["n_objects"][b])
        np.testing.assert_allclose(means.detach().cpu().numpy()[b, :n], refs[b][1], atol=1e-5)
    if bs > 1:
        # q-regulariser couples the images through its denominator: check the loss against the explicit combination
        emb64 = d["emb"].astype(np.float64)
        var = 0.0
        qnum, nfg = 0.0, 0
        for b in range(bs):
            lab = d["labels"][b].reshape(-1)
            xs = emb64[b].reshape(C, -1).T
            fg = lab != 255
            mu = refs[b][1]
            dist = np.linalg.norm(xs[fg] - mu[lab[fg]], axis=1)
            var += (np.clip(dist - 0.5, 0, None) ** 2).sum() / fg.sum()
            l = np.where(fg, np.linalg.norm(xs, axis=1), 0.0)
            qnum += ((l - 1.0) ** 2).sum()
            nfg += int(fg.sum())
        l_ref = var / bs + 0.005 * qnum / nfg
        assert abs(float(loss) - l_ref) <= 1e-4 * abs(l_ref)


@pytest.mark.parametrize("C", [8, 32])
@pytest.mark.parametrize("k", [1, 8, 128])
@pytest.mark.parametrize("HW", [256, 512])
def test_clustering_sweep_bit_exact(cuda, C, k, HW):
    from isa_b200 import clustering
    if HW == 512 and k == 8:
        pytest.skip("the 512^2 column of the sweep runs at the two extreme cluster counts")
    d = synth.batch(C + k, 1, C, HW, HW, min(max(k, 2), 254), n_min=min(max(k, 2), 254), n_max=min(max(k, 2), 254), pull=0.75)
    lab = d["labels"][0]
    sem = np.stack([(lab == 255).astype(np.float32), (lab != 255).astype(np.float32)])
    fg, X = KM.gather_foreground(sem, d["emb"][0])
    budget = dict(seed=3, n_init=2, max_iter=4)
    o = KM.kmeans_oracle(X, k, **budget)
    labels, res = clustering.kmeans_fit_predict(torch.tensor(X, device=cuda), k, **budget)
    assert np.array_equal(res.seed_idx.cpu().numpy(), o["seed_idx"])
    assert np.array_equal(res.n_iter.cpu().numpy(), o["n_iter"])
    assert np.array_equal(res.inertia.cpu().numpy(), o["inertia"])
    assert np.array_equal(labels.cpu().numpy(), o["labels"])
    assert np.array_equal(res.centers.cpu().numpy(), o["centers"])


@pytest.mark.parametrize("C,K,n_max", [(8, 1, 1), (32, 128, 128), (24, 64, 40)])
def test_loss_sweep_1024(cuda, C, K, n_max):
    """The 1024^2 corner of configs[4]: loss, means and gradient of one image."""
    from isa_b200.losses import DiscriminativeLoss
    d = synth.batch(C + K, 1, C, 1024, 1024, K, n_min=max(n_max // 2, 1), n_max=n_max, pull=0.7)
    x = torch.tensor(d["emb"], device=cuda, requires_grad=True)
    loss, means = DiscriminativeLoss(0.5, 1.5, 2)(x, torch.tensor(d["labels"], device=cuda), torch.tensor(d["n_objects"], device=cuda), K)
    loss.backward()
    n = int(d["n_objects"][0])
    l_ref, mu_ref, g_ref = _loss_ref_label_map(d["emb"][0], d["labels"][0], n)
    assert abs(float(loss) - l_ref) <= 1e-4 * abs(l_ref)
    np.testing.assert_allclose(means.detach().cpu().numpy()[0, :n], mu_ref, atol=1e-5)
    assert _rel(x.grad, g_ref[None]) < 1e-4


@pytest.mark.parametrize("C,k", [(8, 1), (32, 64)])
def test_clustering_sweep_1024_bit_exact(cuda, C, k):
    """~520 k foreground points of a 1024^2 image: bit-exact against the C oracle with a bounded restart / iteration budget."""
    from isa_b200 import clustering
    kk = max(k, 2)
    d = synth.batch(C + k + 1, 1, C, 1024, 1024, kk, n_min=kk, n_max=kk, pull=0.75)
    lab = d["labels"][0]
    sem = np.stack([(lab == 255).astype(np.float32), (lab != 255).astype(np.float32)])
    fg, X = KM.gather_foreground(sem, d["emb"][0])
    budget = dict(seed=5, n_init=2, max_iter=3)
    o = KM.kmeans_oracle(X, k, **budget)
    labels, res = clustering.kmeans_fit_predict(torch.tensor(X, device=cuda), k, **budget)
    assert np.array_equal(res.seed_idx.cpu().numpy(), o["seed_idx"])
    assert np.array_equal(res.n_iter.cpu().numpy(), o["n_iter"])
    assert np.array_equal(res.inertia.cpu().numpy(), o["inertia"])
    assert np.array_equal(labels.cpu().numpy(), o["labels"])
