"""CPU: the restatements in oracle/spatial_ref.py (and the product's host-side position encodings) against golden
vectors produced by the reference's own classes (tests/golden/spatial.npz, make_golden_spatial.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import spatial_ref as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_spatial import CASES, D_K, D_V, N_HEAD, inputs  # noqa: E402


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "spatial.npz"))


def _close(a, b, tol=2e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30), float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_masked_softmaxes(G, case):
    name, seed, b, c, h, w, n, n_max = case
    base, sem, ins, q = inputs(case)
    # HardAttentionLayer tail (utils.py:648-652): instance channels >= n_objects are empty -> zeros
    org = torch.tensor(G[name + "_hard_org"]).view(b, h * w)
    y = O.masked_softmax_hw_ref(org, torch.tensor(ins).view(b, n, h * w), None, True)
    _close(y.view(b, n, h, w), G[name + "_hard_split"])
    assert float(y[:, n_max:].abs().max()) == 0.0
    # SpatialAttentionLayer softmax (utils.py:507-512)
    for tag in ("sp", "sp2"):
        logits = torch.tensor(G["%s_%s_beta_logits" % (name, tag)]).view(b, h * w)
        m = torch.tensor(sem).view(b, 1, h * w)
        beta = O.masked_softmax_hw_ref(logits, m, m.sum(2), False)
        _close(beta.view(b, 1, h, w), G["%s_%s_beta" % (name, tag)])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_squeeze_excite_and_readout(G, case):
    name = case[0]
    base, sem, ins, q = inputs(case)
    W = lambda k: torch.tensor(G[name + "_se_w_" + k])  # noqa: E731
    y = O.squeeze_excite_ref(torch.tensor(base), W("fc.0.weight"), W("fc.0.bias"), W("fc.2.weight"), W("fc.2.bias"))
    _close(y, G[name + "_se_y"])
    _close(O.readout_ref(torch.tensor(q), torch.tensor(base)), G[name + "_ro_y"])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("dil", [1, 3])
def test_local_attention_block(G, case, dil):
    name, seed, b, c, h, w, n, n_max = case
    base, sem, ins, q = inputs(case)
    tag = "%s_la%d" % (name, dil)
    W = lambda k: torch.tensor(G[tag + "_w_" + k])  # noqa: E731
    x = torch.tensor(base)
    nomask = 1 - torch.tensor(sem)
    xh = x.view(b * N_HEAD, c // N_HEAD, h, w)
    QK = F.conv2d(xh, W("attention.qk_w.weight"), W("attention.qk_w.bias"))
    V = F.conv2d(xh, W("attention.v_w.weight"), W("attention.v_w.bias"))
    Q, K = QK[:, :D_K], QK[:, D_K:]
    att = O.local_attention_ref(Q, K, V, nomask.repeat(N_HEAD, 1, 1, 1), dil, (c // N_HEAD) ** -0.5)
    att = att.reshape(b, D_V * N_HEAD, h, w)
    out = F.instance_norm(F.conv2d(att, W("attention.fc.weight"), W("attention.fc.bias")) + x)
    ff = F.conv2d(F.leaky_relu(F.conv2d(out, W("w1.weight"), W("w1.bias")), 0.01), W("w2.weight"), W("w2.bias"))
    _close(F.instance_norm(ff + out), G[tag + "_y"], 5e-5)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_mask_bn_inside_hard_attention(G, case):
    name = case[0]
    base, sem, ins, q = inputs(case)
    W = lambda k: torch.tensor(G[name + "_hard_w_" + k])  # noqa: E731
    S = F.avg_pool2d(torch.tensor(base), 3, 1, 1)
    e = F.conv2d(torch.tanh(F.conv2d(S, W("l1.weight"), W("l1.bias"))), W("attend_fc.1.weight"), W("attend_fc.1.bias"), padding=1)
    # the golden state_dict was taken before the forward pass, i.e. it holds the initial affine parameters
    e = O.mask_bn_train_ref(e, torch.tensor(sem), W("bn.weight"), W("bn.bias"))
    e = F.avg_pool2d(e, 3, 1, 1) * torch.tensor(sem)
    _close(e, G[name + "_hard_org"], 5e-5)


def test_position_encodings(G):
    from isa_b200.spatial_attention import make_position_encoding, position_encoding_2d
    np.testing.assert_array_equal(make_position_encoding(np, 2, 37, 24), G["pe_1d"])
    np.testing.assert_array_equal(position_encoding_2d(24, 9, 14).numpy(), G["pe_2d"])
