"""GPU: SBD / best dice / object counts on the device (csrc/eval_ops.cu through the C-ABI) against the reference's own
evaluate.py loops (oracle/evaluate_ref.py, verbatim) -- float64 equality -- including the Cityscapes-shaped size."""
import numpy as np
import pytest
import torch

from isa_b200 import metrics, synth
from oracle import evaluate_ref as E

pytestmark = pytest.mark.gpu


def _pair(seed, H, W, n_gt, n_pr):
    rs = np.random.RandomState(seed)
    gt = synth.label_map(rs, H, W, n_gt).astype(np.int64) + 1
    gt[gt == 256] = 0
    pr = synth.label_map(rs, H, W, n_pr, fg_frac=0.4).astype(np.int64) + 1
    pr[pr == 256] = 0
    return gt.astype(np.uint8), pr.astype(np.uint8)


@pytest.mark.parametrize("seed,H,W,n_gt,n_pr", [(0, 96, 80, 7, 9), (1, 530, 500, 16, 16), (2, 64, 64, 1, 3), (3, 33, 47, 23, 2)])
def test_sbd_device_equals_reference_loops(cuda, seed, H, W, n_gt, n_pr):
    gt, pr = _pair(seed, H, W, n_gt, n_pr)
    out = metrics.sbd_device(torch.tensor(gt, device=cuda), torch.tensor(pr, device=cuda)).cpu().numpy()[0]
    assert out[0] == E.calc_sbd(gt, pr)
    assert out[1] == E.calc_bd(gt, pr) and out[2] == E.calc_bd(pr, gt)
    assert abs(out[3] - out[4]) == E.calc_dic(len(np.unique(gt)) - 1, len(np.unique(pr)) - 1)
    same = metrics.sbd_device(torch.tensor(gt, device=cuda), torch.tensor(gt, device=cuda)).cpu().numpy()[0]
    assert same[0] == 1.0


def test_sbd_device_batch_and_empty_images(cuda):
    pairs = [_pair(10 + i, 48, 56, 3 + i, 5) for i in range(4)]
    gt = np.stack([p[0] for p in pairs])
    pr = np.stack([p[1] for p in pairs])
    pr[3] = 0                                         # no predicted object: the reference raises, the device reports NaN
    out = metrics.sbd_device(torch.tensor(gt, device=cuda), torch.tensor(pr, device=cuda)).cpu().numpy()
    for i in range(3):
        assert out[i, 0] == metrics.calc_sbd(gt[i], pr[i])
    assert np.isnan(out[3, 0]) and out[3, 4] == 0
    with pytest.raises(ValueError):
        metrics.calc_bd(gt[3], pr[3])


def test_sbd_device_cityscapes_shape(cuda):
    """1024 x 2048, 64 instances: the numpy loops are O(n_gt * n_pred * HW); compare against the contingency-table host
    version (itself pinned to the reference loops in tests/test_metrics.py)."""
    gt, pr = _pair(5, 1024, 2048, 64, 60)
    out = metrics.sbd_device(torch.tensor(gt, device=cuda), torch.tensor(pr, device=cuda)).cpu().numpy()[0]
    assert out[0] == metrics.calc_sbd(gt, pr)
    assert (out[3], out[4]) == (64, 60)
