"""TEST INFRASTRUCTURE ONLY -- restatement (torch, float64, autograd for the gradient) of the semantic-head losses
the training step sums next to the discriminative loss:
  dice_coefficient  /root/reference/code/lib/losses/dice.py:10-51
  dice_loss         /root/reference/code/lib/losses/dice.py:54-89
  cross entropy     torch.nn.CrossEntropyLoss(class_weights) over argmax of the one-hot, /root/reference/code/lib/model.py:255-263
Pinned by tests/golden/seg_losses.npz (outputs + input gradients of the reference's own dice_loss, made by
tests/golden/make_golden_seg.py) and, when /root/reference is mounted, live against the reference file."""
import torch
import torch.nn.functional as F


def dice_coefficient(input, target, smooth=1.0, time=2):
    probs = F.softmax(input, dim=1)                          # dice.py:24
    target_f = target.to(probs.dtype)
    num = (probs * target_f).sum(dim=3).sum(dim=2)           # dice.py:27-31
    den1 = (probs if time == 1 else probs * probs).sum(dim=3).sum(dim=2)        # dice.py:32-39
    den2 = (target_f if time == 1 else target_f * target_f).sum(dim=3).sum(dim=2)  # dice.py:40-47
    return (2 * num + smooth) / (den1 + den2 + smooth)       # dice.py:49


def dice_loss(input, target, optimize_bg=False, weight=None, smooth=1.0, time=2):
    dice = dice_coefficient(input, target, smooth=smooth, time=time)
    if not optimize_bg:
        dice = dice[:, 1:]                                   # dice.py:68-70
    if weight is not None:
        if not optimize_bg:
            weight = weight[1:]
        weight = weight.size(0) * weight / weight.sum()      # dice.py:72-76
        dice = dice * weight
    return (1 - dice.mean(1)).mean()                         # dice.py:79-87


def seg_losses(logits, target_one_hot, class_weights=None, optimize_bg=False, smooth=1.0, time=1, g_ce=1.0, g_dice=1.0):
    """-> dict(ce, dice, grad) in float64; grad = d(g_ce * ce + g_dice * dice) / d logits."""
    z = torch.as_tensor(logits).double().clone().requires_grad_(True)
    t = torch.as_tensor(target_one_hot)
    w = None if class_weights is None else torch.as_tensor(class_weights).double()
    ce = F.cross_entropy(z, t.max(1)[1], weight=w)           # model.py:256-262
    dice = dice_loss(z, t.double(), optimize_bg, w, smooth, time)
    (g_ce * ce + g_dice * dice).backward()
    return dict(ce=float(ce), dice=float(dice), grad=z.grad.numpy())
