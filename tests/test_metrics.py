"""CPU: isa_b200.metrics (contingency-table SBD) == the reference's evaluate.py loops (oracle/evaluate_ref.py)."""
import numpy as np
import pytest

from isa_b200 import metrics, synth
from oracle import evaluate_ref as E


@pytest.mark.parametrize("seed", range(5))
def test_sbd_dic_identical_to_reference_functions(seed):
    rs = np.random.RandomState(seed)
    gt = synth.label_map(rs, 96, 80, 7).astype(np.int64) + 1
    gt[gt == 256] = 0
    pred = synth.label_map(rs, 96, 80, 9).astype(np.int64) + 1
    pred[pred == 256] = 0
    assert metrics.calc_sbd(gt, pred) == E.calc_sbd(gt, pred)
    assert metrics.calc_bd(gt, pred) == E.calc_bd(gt, pred)
    assert metrics.calc_dice(gt > 0, pred > 0) == E.calc_dice(gt > 0, pred > 0)
    assert metrics.calc_dic(7, 9) == E.calc_dic(7, 9)
    assert metrics.calc_sbd(gt, gt) == 1.0
