"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's dense attention.

  sdpa_ref      /root/reference/code/lib/archs/modules/utils.py:314-329 (ScaledDotProductAttention.forward)
  MHARef        /root/reference/code/lib/archs/modules/utils.py:167-225 (MultiHeadAttention), eval-mode dropout
Pinned against the reference classes themselves: tests/golden/attention.npz is produced by importing
utils.py from /root/reference (tests/golden/make_golden_attn.py).
"""
import numpy as np
import torch
import torch.nn as nn


def sdpa_ref(q, k, v, temperature, mask=None):
    attn = torch.bmm(q, k.transpose(1, 2)) / temperature
    if mask is not None:
        attn = attn.masked_fill(mask.bool(), -np.inf)
    attn = torch.softmax(attn, dim=2)
    return torch.bmm(attn, v), attn


class MHARef(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)

    def forward(self, q, k, v, mask=None):
        n_head, d_k, d_v = self.n_head, self.d_k, self.d_v
        b, lq, _ = q.shape
        lk = k.shape[1]
        residual = q
        q = self.w_qs(q).view(b, lq, n_head, d_k).permute(2, 0, 1, 3).reshape(-1, lq, d_k)
        k = self.w_ks(k).view(b, lk, n_head, d_k).permute(2, 0, 1, 3).reshape(-1, lk, d_k)
        v = self.w_vs(v).view(b, lk, n_head, d_v).permute(2, 0, 1, 3).reshape(-1, lk, d_v)
        if mask is not None:
            mask = mask.repeat(n_head, 1, 1)
        out, attn = sdpa_ref(q, k, v, float(np.power(d_k, 0.5)), mask)
        out = out.view(n_head, b, lq, d_v).permute(1, 2, 0, 3).reshape(b, lq, -1)
        out = self.layer_norm(self.fc(out) + residual)
        return out, attn
