"""Semantic-head losses used next to the discriminative loss in the training step
(/root/reference/code/lib/losses/dice.py:10-89, lib/model.py:255-269): cross entropy + Dice over the
(b, n_classes, H, W) logits, fused into ONE forward and ONE backward kernel (csrc/seg_losses.cu; SURVEY.md
section 8f rank 2).  The target is the reference collate's one-hot (b, n_classes, H, W) float / int64 / uint8
(distilled once to a 1 B/pixel class map on the device) or, directly, a (b, H, W) uint8 class map.

    ce, dice = SegLosses(class_weights, optimize_bg, smooth)(logits, target, time=1)
    DiceLoss(optimize_bg, weight, smooth)(logits, target, time=2)          # dice.py:54-89 as a module

dice.py's host-side np.unique asserts (a device sync per step) are dropped; a one-hot target is assumed, as the
reference's collate guarantees (lib/dataset.py:354-376).
"""
import torch
from torch.nn.modules.loss import _Loss

from . import _lib

_KIND = {torch.float32: 1, torch.int64: 2, torch.uint8: 3, torch.bool: 3}


def class_map(target):
    """(b,H,W) uint8 class map as is; (b,nc,H,W) one-hot -> uint8 class map (first maximum, like Tensor.max(1)[1])."""
    _lib.require_cuda(target, "target")
    if target.dim() == 3:
        if target.dtype != torch.uint8:
            raise TypeError("a (b,H,W) class-map target must be uint8")
        return target.contiguous()
    if target.dim() != 4 or target.dtype not in _KIND:
        raise TypeError("target must be (b,nc,H,W) float32/int64/uint8 one-hot or a (b,H,W) uint8 class map")
    lib = _lib.load()
    t = target.contiguous()
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    b, nc, H, W = t.shape
    out = torch.empty(b, H, W, device=t.device, dtype=torch.uint8)
    rc = lib.isa_onehot_argmax(_lib.ptr(t), _KIND[target.dtype], b, nc, H * W, _lib.ptr(out), _lib.stream_ptr(t.device))
    _lib.check(rc, "isa_onehot_argmax")
    return out


class _SegLossFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, logits, cmap, class_w, dice_time, smooth, optimize_bg):
        lib = _lib.load()
        _lib.require_cuda(logits, "input")
        if logits.dtype != torch.float32:
            raise TypeError("the semantic losses compute in fp32; got %s" % logits.dtype)
        z = logits.contiguous()
        b, nc, H, W = z.shape
        if cmap.shape != (b, H, W):
            raise ValueError("target %s does not match the logits %s" % (tuple(cmap.shape), tuple(z.shape)))
        w = None if class_w is None else class_w.to(device=z.device, dtype=torch.float32).contiguous()
        ws_bytes = lib.isa_seg_losses_workspace_bytes(b, nc, H * W)
        ws = torch.empty(ws_bytes, device=z.device, dtype=torch.uint8)
        out = torch.empty(2, device=z.device, dtype=torch.float32)
        rc = lib.isa_seg_losses_fwd(_lib.ptr(z), _lib.ptr(cmap), _lib.ptr(w), b, nc, H * W, int(dice_time), float(smooth),
                                    int(bool(optimize_bg)), _lib.ptr(out), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(z.device))
        _lib.check(rc, "isa_seg_losses_fwd")
        ctx.save_for_backward(z, cmap, w, ws)
        ctx.cfg = (b, nc, H * W, int(dice_time))
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_ce, g_dice):
        lib = _lib.load()
        z, cmap, w, ws = ctx.saved_tensors
        b, nc, HW, dice_time = ctx.cfg
        gce = None if g_ce is None else g_ce.reshape(1).to(torch.float32).contiguous()
        gd = None if g_dice is None else g_dice.reshape(1).to(torch.float32).contiguous()
        grad = torch.empty_like(z)
        rc = lib.isa_seg_losses_bwd(_lib.ptr(z), _lib.ptr(cmap), _lib.ptr(w), b, nc, HW, dice_time, _lib.ptr(gce), _lib.ptr(gd),
                                    _lib.ptr(grad), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(z.device))
        _lib.check(rc, "isa_seg_losses_bwd")
        return grad, None, None, None, None, None


class SegLosses(_Loss):
    """(cross entropy, Dice loss) of the semantic head in one pass; both are means as in the reference
    (CrossEntropyLoss(weight) and dice_loss(size_average=True, reduce=True))."""

    def __init__(self, class_weights=None, optimize_bg=False, smooth=1.0):
        super(SegLosses, self).__init__()
        assert smooth > 0, 'Smooth must be greater than 0.'           # dice.py:22
        self.weight = class_weights
        self.optimize_bg = bool(optimize_bg)
        self.smooth = float(smooth)

    def forward(self, input, target, time=2):
        return _SegLossFn.apply(input, class_map(target), self.weight, time, self.smooth, self.optimize_bg)


class DiceLoss(_Loss):
    """dice.py:54-89 (dice_loss) as a module, on the fused kernel (the cross-entropy half gets no gradient)."""

    def __init__(self, optimize_bg=False, weight=None, smooth=1.0, size_average=True, reduce=True):
        super(DiceLoss, self).__init__()
        if not (size_average and reduce):
            raise NotImplementedError("DiceLoss: only the mean reduction the reference's Model uses is implemented")
        self.fused = SegLosses(weight, optimize_bg, smooth)

    def forward(self, input, target, time=2):
        return self.fused(input, target, time=time)[1]
