"""GPU: the sm_100a discriminative-loss kernels (through the C-ABI) against the numpy oracle
and against golden vectors produced by the reference itself.  Tolerance: north_star's
1e-3 relative for fp32 (we assert 1e-4: the kernels accumulate in fp32 / fp64 counters)."""
import os
import sys

import numpy as np
import pytest
import torch

from isa_b200 import synth
from oracle import disc_loss as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import DISC_CASES  # noqa: E402

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _run(dev, emb, tgt, nobj, K, norm=2, terms=None, normalize=True, grad_means=None):
    from isa_b200.losses import DiscriminativeLoss, SHIPPED_TERMS
    x = torch.tensor(emb, device=dev, requires_grad=True)
    t = torch.tensor(tgt, device=dev)
    n = torch.tensor(nobj, device=dev)
    crit = DiscriminativeLoss(0.5, 1.5, norm, terms=terms or SHIPPED_TERMS, normalize_means=normalize)
    loss, means = crit(x, t, n, K)
    total = loss
    if grad_means is not None:
        total = loss + (means * torch.tensor(grad_means, device=dev)).sum()
    total.backward()
    torch.cuda.synchronize()
    return float(loss), means.detach().cpu().numpy(), x.grad.cpu().numpy(), crit.last_terms.cpu().numpy()


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() <= rtol * scale


@pytest.mark.parametrize("case", DISC_CASES, ids=[c[0] for c in DISC_CASES])
@pytest.mark.parametrize("kind", ["label", "f32", "i64", "u8"])
def test_against_reference_golden(cuda, golden_dir, case, kind):
    name, seed, bs, C, H, W, K, norm, kw = case
    g = np.load(os.path.join(golden_dir, "disc_loss.npz"))
    d = synth.batch(seed, bs, C, H, W, K, **kw)
    tgt = {"label": d["labels"], "f32": synth.onehot(d["labels"], K), "i64": synth.onehot(d["labels"], K, np.int64),
           "u8": synth.onehot(d["labels"], K, np.uint8)}[kind]
    loss, means, grad, terms = _run(cuda, d["emb"], tgt, d["n_objects"], K, norm, grad_means=g[name + "_gm"])
    assert abs(loss - float(g[name + "_loss"])) <= RTOL * abs(float(g[name + "_loss"]))
    np.testing.assert_allclose(means, g[name + "_means"], atol=1e-5)
    assert _close(grad, g[name + "_grad"])
    assert _close(terms[[0, 3]], g[name + "_terms"][[0, 3]])


@pytest.mark.parametrize("C,K,H,W,bs", [(8, 1, 17, 19, 1), (16, 4, 64, 96, 3), (24, 32, 256, 256, 2), (32, 64, 96, 160, 2),
                                        (48, 8, 32, 32, 2), (24, 32, 40, 40, 400)])
@pytest.mark.parametrize("norm", [1, 2])
def test_against_oracle_shapes(cuda, C, K, H, W, bs, norm):
    if bs == 400 and norm == 1:
        pytest.skip("one norm is enough for the bs > co-resident-CTA path")
    d = synth.batch(C * 1000 + K, bs, C, H, W, K, n_min=1, n_max=min(K, 23))
    o = O.discriminative_loss(d["emb"], d["labels"], d["n_objects"], K, 0.5, 1.5, norm, want_grad=True)
    loss, means, grad, terms = _run(cuda, d["emb"], d["labels"], d["n_objects"], K, norm)
    assert abs(loss - float(o["loss"])) <= RTOL * abs(float(o["loss"]))
    np.testing.assert_allclose(means, o["means"], atol=1e-5)
    assert _close(grad, o["grad"])


@pytest.mark.parametrize("norm", [1, 2])
@pytest.mark.parametrize("normalize", [True, False])
def test_all_terms_paper_form(cuda, norm, normalize):
    terms = (1.0, 1.0, 0.001, 0.005)
    d = synth.batch(77, 3, 16, 48, 48, 8, n_min=1, n_max=8, pull=0.3)
    gm = np.random.RandomState(3).standard_normal((3, 8, 16)).astype(np.float32) * 0.01
    o = O.discriminative_loss(d["emb"], d["labels"], d["n_objects"], 8, 0.5, 1.5, norm, terms=terms,
                              normalize_means=normalize, want_grad=True, grad_means=gm)
    loss, means, grad, t = _run(cuda, d["emb"], d["labels"], d["n_objects"], 8, norm, terms=terms,
                                normalize=normalize, grad_means=gm)
    assert abs(loss - float(o["loss"])) <= RTOL * abs(float(o["loss"]))
    assert _close(t, o["terms"])
    np.testing.assert_allclose(means, o["means"], atol=1e-5)
    assert _close(grad, o["grad"])


def test_soft_and_overlapping_masks(cuda):
    rs = np.random.RandomState(7)
    bs, C, H, W, K = 2, 12, 20, 24, 5
    emb = rs.standard_normal((bs, C, H, W)).astype(np.float32)
    tgt = (rs.uniform(size=(bs, K, H, W)) * (rs.uniform(size=(bs, K, H, W)) > 0.5)).astype(np.float32)
    nobj = np.array([3, 5], dtype=np.int32)
    o = O.discriminative_loss(emb, tgt, nobj, K, 0.5, 1.5, 2, want_grad=True)
    loss, means, grad, _ = _run(cuda, emb, tgt, nobj, K)
    assert abs(loss - float(o["loss"])) <= RTOL * abs(float(o["loss"]))
    np.testing.assert_allclose(means, o["means"], atol=1e-5)
    assert _close(grad, o["grad"])


def test_edge_cases(cuda):
    # instance id below n_objects without pixels -> NaN like the reference; labels >= n_objects are
    # foreground for the q-regulariser only; all-background image -> 0/0.
    d = synth.batch(5, 2, 8, 16, 16, 4, n_min=2, n_max=2)
    n = np.array([3, 2], dtype=np.int32)
    loss, means, grad, _ = _run(cuda, d["emb"], d["labels"], n, 4)
    assert np.isnan(loss)
    n = np.array([1, 2], dtype=np.int32)
    o = O.discriminative_loss(d["emb"], d["labels"], n, 4, 0.5, 1.5, 2, want_grad=True)
    loss, means, grad, _ = _run(cuda, d["emb"], d["labels"], n, 4)
    assert abs(loss - float(o["loss"])) <= RTOL * abs(float(o["loss"]))
    assert _close(grad, o["grad"])
    assert np.all(means[0, 1:] == 0) and np.all(means[1, 2:] == 0)


def test_onehot_to_labels(cuda):
    from isa_b200.losses import onehot_to_labels
    d = synth.batch(9, 3, 8, 33, 47, 16)
    for dt in (np.float32, np.int64, np.uint8):
        lab, flag = onehot_to_labels(torch.tensor(synth.onehot(d["labels"], 16, dt), device=cuda))
        assert np.array_equal(lab.cpu().numpy(), d["labels"]) and int(flag) == 0
    soft = synth.onehot(d["labels"], 16) * 0.5
    _, flag = onehot_to_labels(torch.tensor(soft, device=cuda))
    assert int(flag) == 1


def test_errors_are_loud(cuda):
    from isa_b200 import _lib
    from isa_b200.losses import DiscriminativeLoss
    crit = DiscriminativeLoss(0.5, 1.5, 2)
    with pytest.raises(_lib.IsaError):
        crit(torch.zeros(1, 4, 8, 8), torch.zeros(1, 2, 8, 8), torch.tensor([1]), 2)  # CPU tensors
    with pytest.raises(_lib.IsaError):
        crit(torch.zeros(1, 80, 8, 8, device=cuda), torch.zeros(1, 8, 8, device=cuda, dtype=torch.uint8), torch.tensor([1]), 2)


def test_full_size_properties(cuda):
    """BASELINE config size (bs=16, C=24, K=32, 256x256): size-independent properties instead of
    the (slow) dense oracle: label-map and one-hot targets agree; loss is invariant to a
    permutation of instance ids; grad of a constant upstream scale is linear."""
    from isa_b200.losses import DiscriminativeLoss
    d = synth.batch(123, 16, 24, 256, 256, 32)
    crit = DiscriminativeLoss(0.5, 1.5, 2)
    x = torch.tensor(d["emb"], device=cuda, requires_grad=True)
    lab = torch.tensor(d["labels"], device=cuda)
    n = torch.tensor(d["n_objects"], device=cuda)
    l1, m1 = crit(x, lab, n, 32)
    (g1,) = torch.autograd.grad(l1, x)
    oh = torch.tensor(synth.onehot(d["labels"], 32, np.int64), device=cuda)
    l2, m2 = crit(x, oh, n, 32)
    (g2,) = torch.autograd.grad(3.0 * l2, x)
    assert abs(float(l1) - float(l2)) <= 1e-5 * abs(float(l1))
    assert torch.allclose(m1, m2, atol=1e-5)
    assert float((g2 - 3.0 * g1).abs().max()) <= 1e-4 * float(g1.abs().max()) * 3
    # permute instance ids inside each image
    labp = d["labels"].copy()
    for b in range(16):
        nb = int(d["n_objects"][b])
        perm = np.random.RandomState(b).permutation(nb)
        m = labp[b] != 255
        labp[b][m] = perm[labp[b][m]]
    l3, _ = crit(x, torch.tensor(labp, device=cuda), n, 32)
    assert abs(float(l1) - float(l3)) <= 1e-5 * abs(float(l1))
