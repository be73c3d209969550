"""GPU: the fused cross entropy + Dice kernels (csrc/seg_losses.cu) against the oracle restatement and the golden vectors
of the reference's own dice_loss; every target format; bitwise run-to-run determinism; full CVPPP batch shape."""
import os

import numpy as np
import pytest
import torch

from oracle import seg_losses_ref as O
from tests.golden.make_golden_seg import SEG_CASES, seg_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "seg_losses.npz"))


def _run(cuda, logits, target, w, obg, time, g_ce, g_dice):
    from isa_b200.seg_losses import SegLosses
    z = torch.tensor(logits, device=cuda, requires_grad=True)
    wt = None if w is None else torch.tensor(w, dtype=torch.float32, device=cuda)
    ce, dice = SegLosses(wt, obg, 1.0)(z, target, time=time)
    (g_ce * ce + g_dice * dice).backward()
    return float(ce), float(dice), z.grad.cpu().numpy()


@pytest.mark.parametrize("case", SEG_CASES, ids=[c[0] for c in SEG_CASES])
def test_golden_and_oracle(cuda, golden, case):
    name, seed, bs, nc, H, W, w, obg, time = case
    logits, cls, onehot = seg_case(case)
    ce, dice, grad = _run(cuda, logits, torch.tensor(cls, device=cuda), w, obg, time, 0.7, 1.3)
    assert abs(ce - float(golden[name + "_ce"])) < 1e-5 * abs(ce)          # fp32 tolerance of north_star: 1e-3
    assert abs(dice - float(golden[name + "_dice"])) < 1e-5
    g = golden[name + "_grad"]
    assert np.abs(grad - g).max() < 1e-4 * np.abs(g).max()
    o = O.seg_losses(logits, onehot, w, obg, 1.0, time, 0.7, 1.3)
    assert abs(ce - o["ce"]) < 1e-5 * abs(o["ce"]) and abs(dice - o["dice"]) < 1e-5
    assert np.abs(grad - o["grad"]).max() < 1e-4 * np.abs(o["grad"]).max()


@pytest.mark.parametrize("dtype", [torch.int64, torch.float32, torch.uint8])
def test_one_hot_formats_equal_class_map(cuda, dtype):
    case = ("fmt", 5, 2, 2, 30, 50, [0.5, 1.5], False, 1)
    logits, cls, onehot = seg_case(case)
    a = _run(cuda, logits, torch.tensor(cls, device=cuda), case[6], False, 1, 1.0, 1.0)
    b = _run(cuda, logits, torch.tensor(onehot, device=cuda).to(dtype), case[6], False, 1, 1.0, 1.0)
    assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])


def test_dice_module_and_deterministic_at_batch_shape(cuda):
    """(16,2,256,256): the training step's shape; two runs must agree bit for bit (fixed-order partial sums)."""
    from isa_b200.seg_losses import DiceLoss
    case = ("full", 6, 16, 2, 256, 256, None, False, 1)
    logits, cls, onehot = seg_case(case)
    r1 = _run(cuda, logits, torch.tensor(cls, device=cuda), None, False, 1, 1.0, 1.0)
    r2 = _run(cuda, logits, torch.tensor(cls, device=cuda), None, False, 1, 1.0, 1.0)
    assert r1[0] == r2[0] and r1[1] == r2[1] and np.array_equal(r1[2], r2[2])
    o = O.seg_losses(logits, onehot, None, False, 1.0, 1, 1.0, 1.0)
    assert abs(r1[0] - o["ce"]) < 1e-5 * o["ce"] and abs(r1[1] - o["dice"]) < 1e-5
    assert np.abs(r1[2] - o["grad"]).max() < 1e-4 * np.abs(o["grad"]).max()
    z = torch.tensor(logits, device=cuda)
    d = DiceLoss()(z, torch.tensor(onehot, device=cuda), time=1)
    assert abs(float(d) - o["dice"]) < 1e-5


def test_rejects_cpu_and_bad_shapes(cuda):
    from isa_b200 import _lib
    from isa_b200.seg_losses import SegLosses
    with pytest.raises(_lib.IsaError):
        SegLosses()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.uint8))
    with pytest.raises(ValueError):
        SegLosses()(torch.zeros(1, 2, 4, 4, device=cuda), torch.zeros(1, 5, 4, dtype=torch.uint8, device=cuda))
