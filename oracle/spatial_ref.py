"""TEST INFRASTRUCTURE ONLY -- CPU (torch, any dtype) restatements of the reference's spatial / channel /
local attention arithmetic, each citing the reference lines it follows
(/root/reference/code/lib/archs/modules/utils.py).  Pinned against the reference classes themselves:
tests/golden/spatial.npz is produced by importing utils.py from /root/reference
(tests/golden/make_golden_spatial.py) and tests/test_oracle_spatial.py checks these functions against it.
"""
import numpy as np
import torch
import torch.nn.functional as F


def masked_softmax_hw_ref(x, mask, scale=None, nan_to_zero=False):
    """x (B,HW), mask (B,K,HW) 0/1 -> (B,K,HW).
    utils.py:507-512 (SpatialAttentionLayer: masked_fill(1-y, -inf), softmax over HW, * sum(y); NaN kept) and
    utils.py:648-652 (HardAttentionLayer: expand to n instances, masked_fill, softmax, NaN -> 0)."""
    B, K, HW = mask.shape
    e = x.view(B, 1, HW).expand(B, K, HW).masked_fill(mask == 0, -np.inf)
    y = torch.softmax(e, dim=2)
    if nan_to_zero:
        y = torch.where(torch.isnan(y), torch.zeros_like(y), y)
    if scale is not None:
        y = y * scale.view(B, K, 1)
    return y


def squeeze_excite_ref(x, w1, b1, w2, b2, multiply=True):
    """utils.py:413-420 AttentionLayer.forward"""
    b, c, _, _ = x.shape
    y = x.mean(dim=(2, 3))
    y = torch.sigmoid(F.linear(torch.relu(F.linear(y, w1, b1)), w2, b2)).view(b, c, 1, 1)
    return x * y if multiply else y


def readout_ref(q, enc):
    """utils.py:59-69 Decoder.forward: sigmoid(bmm(q.unsqueeze(1), enc.view(b,c,-1))).squeeze(1)"""
    b, c = enc.shape[:2]
    return torch.sigmoid(torch.bmm(q.unsqueeze(1), enc.reshape(b, c, -1))).squeeze(1)


def local_attention_ref(Q, K, V, nomask, d, scale):
    """utils.py:279-299 (the part of _ScalePDAttention.forward between the 1x1 projections and `fc`).
    Q,K (Bh,dk,h,w), V (Bh,dv,h,w), nomask (Bh,1,h,w) non-zero = excluded (already tiled to Bh) or None."""
    Bh, dk, h, w = Q.shape
    dv = V.shape[1]
    pad = (d, d, d, d)
    Kp, Vp = F.pad(K, pad), F.pad(V, pad)
    if nomask is None:
        nomask = torch.zeros(Bh, 1, h, w, dtype=Q.dtype)
    Mp = F.pad(nomask, pad)
    sl = [(slice(i // 3 * d, i // 3 * d + h), slice(i % 3 * d, i % 3 * d + w)) for i in range(9)]
    K_ = torch.stack([Kp[:, :, a, b_] for a, b_ in sl], dim=1)      # Bh, 9, dk, h, w
    V_ = torch.stack([Vp[:, :, a, b_] for a, b_ in sl], dim=1)      # Bh, 9, dv, h, w
    M_ = torch.stack([Mp[:, 0, a, b_] for a, b_ in sl], dim=1)      # Bh, 9, h, w
    inner = (K_ * Q.unsqueeze(1)).sum(2) * scale                    # Bh, 9, h, w
    inner = inner.masked_fill(M_ != 0, -np.inf)
    P = torch.softmax(inner, dim=1)
    P = torch.where(torch.isnan(P), torch.zeros_like(P), P)
    return (V_ * P.unsqueeze(2)).sum(1)                             # Bh, dv, h, w


def mask_bn_train_ref(x, mask, weight, bias, eps=1e-5):
    """utils.py:573-586 maskBN.forward, training branch (batch statistics)."""
    b, c, h, w = x.shape
    mask = mask.expand(b, c, h, w)
    mm = mask.reshape(b, -1).sum(1) + 1
    x2, m2 = x.reshape(b, c, -1), mask.reshape(b, c, -1)
    mean = ((x2 * m2).sum(2) / mm[:, None]).mean(0)
    var = ((((x2 - mean[None, :, None]) ** 2) * m2).sum(2) / mm[:, None]).mean(0)
    return (x - mean.view(1, c, 1, 1)) / torch.pow(var.view(1, c, 1, 1) + eps, 0.5) * weight.view(1, c, 1, 1) + bias.view(1, c, 1, 1)
