"""TEST INFRASTRUCTURE ONLY -- the reference's result metrics restated as importable functions.
/root/reference/code/evaluate.py:18-57 (calc_dic, calc_dice, calc_bd, calc_sbd); the reference file
parses argv at import (evaluate.py:6-11) so it cannot be imported directly.  The product's own
evaluate.py (repo root) uses isa_b200.metrics, which tests check against these."""
import numpy as np


def calc_dic(n_objects_gt, n_objects_pred):
    return np.abs(n_objects_gt - n_objects_pred)


def calc_dice(gt_seg, pred_seg):
    nom = 2 * np.sum(gt_seg * pred_seg)
    denom = np.sum(gt_seg) + np.sum(pred_seg)
    return float(nom) / float(denom)


def calc_bd(ins_seg_gt, ins_seg_pred):
    gt_object_idxes = list(set(np.unique(ins_seg_gt)).difference([0]))
    pred_object_idxes = list(set(np.unique(ins_seg_pred)).difference([0]))
    best_dices = []
    for gt_idx in gt_object_idxes:
        _gt_seg = (ins_seg_gt == gt_idx).astype('bool')
        dices = []
        for pred_idx in pred_object_idxes:
            _pred_seg = (ins_seg_pred == pred_idx).astype('bool')
            dices.append(calc_dice(_gt_seg, _pred_seg))
        best_dices.append(np.max(dices))
    return np.mean(best_dices)


def calc_sbd(ins_seg_gt, ins_seg_pred):
    return min(calc_bd(ins_seg_gt, ins_seg_pred), calc_bd(ins_seg_pred, ins_seg_gt))
