"""CPU: the semantic-loss restatement (oracle/seg_losses_ref.py) against golden vectors produced by the reference's own
dice_loss (tests/golden/make_golden_seg.py) and, when the reference tree is mounted, live against the file."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader, seg_losses_ref as O
from tests.golden.make_golden_seg import SEG_CASES, seg_case


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "seg_losses.npz"))


@pytest.mark.parametrize("case", SEG_CASES, ids=[c[0] for c in SEG_CASES])
def test_restatement_matches_reference_golden(golden, case):
    name, seed, bs, nc, H, W, w, obg, time = case
    logits, cls, onehot = seg_case(case)
    o = O.seg_losses(logits, onehot, w, obg, 1.0, time, g_ce=0.7, g_dice=1.3)
    assert abs(o["ce"] - float(golden[name + "_ce"])) < 2e-6 * abs(o["ce"])
    assert abs(o["dice"] - float(golden[name + "_dice"])) < 2e-6
    g = golden[name + "_grad"]
    assert np.abs(o["grad"] - g).max() < 2e-6 * np.abs(g).max() + 1e-10


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_restatement_matches_reference_live():
    dice = ref_loader.load_file("lib/losses/dice.py")
    case = ("live", 9, 2, 2, 17, 23, [0.4, 1.6], False, 1)
    logits, cls, onehot = seg_case(case)
    z = torch.tensor(logits).double().requires_grad_(True)
    w = torch.tensor(case[6]).double()
    d = dice.dice_loss(z, torch.tensor(onehot), optimize_bg=False, weight=w, smooth=1.0, time=1)
    o = O.seg_losses(logits, onehot, case[6], False, 1.0, 1, g_ce=0.0, g_dice=1.0)
    d.backward()
    assert abs(float(d) - o["dice"]) < 1e-12
    assert np.abs(z.grad.numpy() - o["grad"]).max() < 1e-12
