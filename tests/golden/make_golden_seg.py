"""Golden vectors of the semantic-head losses from the REFERENCE's own dice_loss (lib/losses/dice.py, imported
through oracle/ref_loader.py) and torch's CrossEntropyLoss as lib/model.py:255-263 calls it.  Build container only:
    python tests/golden/make_golden_seg.py
Inputs are regenerated from seeds (seg_case) by the tests."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

# name, seed, bs, n_classes, H, W, class weights, optimize_bg, time
SEG_CASES = [
    ("cvppp", 0, 3, 2, 40, 56, None, False, 1),
    ("weighted", 1, 2, 2, 33, 47, [0.3, 1.7], False, 1),
    ("bg", 2, 2, 3, 24, 24, [1.0, 2.0, 0.5], True, 2),
    ("five", 3, 1, 5, 31, 29, None, False, 2),
]


def seg_case(case):
    name, seed, bs, nc, H, W, w, obg, time = case
    rs = np.random.RandomState(seed)
    logits = (2.0 * rs.standard_normal((bs, nc, H, W))).astype(np.float32)
    cls = rs.randint(0, nc, size=(bs, H, W)).astype(np.uint8)
    onehot = np.stack([(cls == c) for c in range(nc)], 1).astype(np.int64)
    return logits, cls, onehot


def main():
    from oracle import ref_loader
    dice = ref_loader.load_file("lib/losses/dice.py")
    out = {}
    for case in SEG_CASES:
        name, seed, bs, nc, H, W, w, obg, time = case
        logits, cls, onehot = seg_case(case)
        z = torch.tensor(logits, requires_grad=True)
        t = torch.tensor(onehot)
        wt = None if w is None else torch.tensor(w, dtype=torch.float32)
        d = dice.dice_loss(z, t, optimize_bg=obg, weight=wt, smooth=1.0, time=time)
        ce = torch.nn.CrossEntropyLoss(wt)(z.permute(0, 2, 3, 1).contiguous().view(-1, nc), t.max(1)[1].view(-1))
        (0.7 * ce + 1.3 * d).backward()
        out[name + "_ce"] = ce.detach().numpy()
        out[name + "_dice"] = d.detach().numpy()
        out[name + "_grad"] = z.grad.numpy()
        print(name, float(ce), float(d))
    np.savez_compressed(os.path.join(HERE, "seg_losses.npz"), **out)


if __name__ == "__main__":
    main()
