"""Semantic-head losses used next to the discriminative loss in the training step
(/root/reference/code/lib/losses/dice.py:10-89, lib/model.py:255-269).  Plain PyTorch: elementwise
+ reduce over (b,2,H,W) logits; SURVEY.md lists fusing them as a later step (section 8f, rank 2)."""
import torch
import torch.nn.functional as F
from torch.nn.modules.loss import _Loss


def dice_coefficient(input, target, smooth=1.0, time=2):
    """dice.py:10-51 without the host-side np.unique asserts (they force a device sync per step)."""
    probs = F.softmax(input, dim=1)
    target_f = target.float()
    num = (probs * target_f).sum(dim=(2, 3))
    den1 = (probs if time == 1 else probs * probs).sum(dim=(2, 3))
    den2 = (target_f if time == 1 else target_f * target_f).sum(dim=(2, 3))
    return (2 * num + smooth) / (den1 + den2 + smooth)


class DiceLoss(_Loss):
    """dice.py:54-89 (dice_loss) as a module."""

    def __init__(self, optimize_bg=False, weight=None, smooth=1.0, size_average=True, reduce=True):
        super(DiceLoss, self).__init__()
        self.optimize_bg = optimize_bg
        self.weight = weight
        self.smooth = smooth
        self.size_average = size_average
        self.reduce = reduce

    def forward(self, input, target, time=2):
        dice = dice_coefficient(input, target, smooth=self.smooth, time=time)
        if not self.optimize_bg:
            dice = dice[:, 1:]
        if self.weight is not None:
            w = self.weight if self.optimize_bg else self.weight[1:]
            w = w.size(0) * w / w.sum()
            dice = dice * w
        loss = 1 - dice.mean(1)
        if not self.reduce:
            return loss
        return loss.mean() if self.size_average else loss.sum()
