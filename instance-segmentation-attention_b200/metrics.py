"""Result metrics of /root/reference/code/evaluate.py:18-57 (|DiC|, dice, best dice, symmetric best
dice) computed from one contingency table instead of an O(n_gt * n_pred * HW) loop.  float64 numpy;
values are identical to the reference's because every dice is the same ratio of integer counts."""
import numpy as np


def calc_dic(n_objects_gt, n_objects_pred):
    return np.abs(n_objects_gt - n_objects_pred)


def calc_dice(gt_seg, pred_seg):
    gt_seg = np.asarray(gt_seg)
    pred_seg = np.asarray(pred_seg)
    nom = 2 * np.sum(gt_seg * pred_seg)
    denom = np.sum(gt_seg) + np.sum(pred_seg)
    return float(nom) / float(denom)


def _contingency(a, b):
    a = np.asarray(a).astype(np.int64).ravel()
    b = np.asarray(b).astype(np.int64).ravel()
    ua, ia = np.unique(a, return_inverse=True)
    ub, ib = np.unique(b, return_inverse=True)
    table = np.zeros((len(ua), len(ub)), dtype=np.int64)
    np.add.at(table, (ia, ib), 1)
    return ua, ub, table


def calc_bd(ins_seg_gt, ins_seg_pred):
    ua, ub, t = _contingency(ins_seg_gt, ins_seg_pred)
    ga, gb = ua != 0, ub != 0
    if not ga.any():
        return np.nan          # np.mean([]) in the reference
    if not gb.any():
        raise ValueError("zero-size array to reduction operation maximum which has no identity")  # np.max([]) in the reference
    size_a = t.sum(1)[ga].astype(np.float64)
    size_b = t.sum(0)[gb].astype(np.float64)
    inter = t[np.ix_(ga, gb)].astype(np.float64)
    dice = 2 * inter / (size_a[:, None] + size_b[None, :])
    return np.mean(dice.max(1))


def calc_sbd(ins_seg_gt, ins_seg_pred):
    return min(calc_bd(ins_seg_gt, ins_seg_pred), calc_bd(ins_seg_pred, ins_seg_gt))


def sbd_device(ins_seg_gt, ins_seg_pred):
    """SBD / best dice / object counts of uint8 label images ON THE DEVICE (csrc/eval_ops.cu): gt, pred (n, H, W) or
    (H, W) uint8 CUDA tensors -> float64 CUDA tensor (n, 5) = [SBD, BD(gt, pred), BD(pred, gt), n_gt, n_pred]; |DiC| is
    |n_gt - n_pred|.  Same values as calc_sbd / calc_bd above (identical ratios of integer counts in float64); NaN where
    calc_bd returns nan or raises.  No CPU fallback."""
    import torch
    from . import _lib
    lib = _lib.load()
    _lib.require_cuda(ins_seg_gt, "ins_seg_gt")
    _lib.require_cuda(ins_seg_pred, "ins_seg_pred")
    if ins_seg_gt.dtype != torch.uint8 or ins_seg_pred.dtype != torch.uint8:
        raise TypeError("sbd_device: label images must be uint8")
    gt = ins_seg_gt.contiguous()
    pr = ins_seg_pred.contiguous()
    if gt.shape != pr.shape or gt.dim() not in (2, 3):
        raise ValueError("sbd_device: gt %s and pred %s must be equal-shaped (H,W) or (n,H,W)" % (tuple(gt.shape), tuple(pr.shape)))
    n = 1 if gt.dim() == 2 else gt.shape[0]
    P = gt.numel() // n
    out = torch.empty(n, 5, device=gt.device, dtype=torch.float64)
    wsb = lib.isa_sbd_workspace_bytes(n)
    ws = torch.empty(wsb, device=gt.device, dtype=torch.uint8)
    rc = lib.isa_sbd(_lib.ptr(gt), _lib.ptr(pr), n, P, _lib.ptr(out), _lib.ptr(ws), wsb, _lib.stream_ptr(gt.device))
    _lib.check(rc, "isa_sbd")
    return out
