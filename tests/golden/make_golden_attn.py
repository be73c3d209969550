"""Golden vectors of the dense attention from the REFERENCE CLASSES themselves
(lib/archs/modules/utils.py MultiHeadAttention / ScaledDotProductAttention, imported from
/root/reference through oracle/ref_loader.py, eval mode).  Run through make_golden.py attn."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
from oracle import ref_loader  # noqa: E402

# name, seed, b, Lq, Lk, n_head, d_model, d_k, d_v, masked
ATTN_CASES = [
    ("cfg", 0, 2, 200, 200, 2, 24, 12, 12, True),
    ("cross", 1, 3, 37, 300, 2, 24, 12, 12, True),
    ("nomask", 2, 1, 129, 129, 1, 16, 16, 8, False),
]


def case_inputs(case):
    name, seed, b, Lq, Lk, n_head, d_model, d_k, d_v, masked = case
    rs = np.random.RandomState(seed)
    q = rs.standard_normal((b, Lq, d_model)).astype(np.float32)
    kv = q if Lq == Lk else rs.standard_normal((b, Lk, d_model)).astype(np.float32)
    mask = None
    if masked:
        mask = (rs.uniform(size=(b, 1, Lk)) < 0.3)
        mask[:, :, 0] = False  # keep at least one key
        mask = mask.astype(np.uint8)
    return q, kv, mask


def main():
    U = ref_loader.attention_utils()
    out = {}
    for case in ATTN_CASES:
        name, seed, b, Lq, Lk, n_head, d_model, d_k, d_v, masked = case
        torch.manual_seed(seed)
        mha = U.MultiHeadAttention(n_head, d_model, d_k, d_v).eval()
        q, kv, mask = case_inputs(case)
        tq = torch.tensor(q, requires_grad=True)
        tkv = torch.tensor(kv, requires_grad=True) if kv is not q else tq
        m = torch.tensor(mask) if mask is not None else None
        y, attn = mha(tq, tkv, tkv, mask=m)
        gy = torch.tensor(np.random.RandomState(100 + seed).standard_normal(y.shape).astype(np.float32))
        y.backward(gy)
        out[name + "_y"] = y.detach().numpy()
        out[name + "_attn_rowsum"] = attn.detach().sum(2).numpy()
        out[name + "_attn_diag"] = attn.detach()[:, : min(Lq, Lk), : min(Lq, Lk)].diagonal(dim1=1, dim2=2).numpy()
        out[name + "_gq"] = tq.grad.numpy()
        out[name + "_gy"] = gy.numpy()
        for k_, v_ in mha.state_dict().items():
            out[name + "_w_" + k_] = v_.numpy()
        for k_, p_ in mha.named_parameters():
            out[name + "_g_" + k_] = p_.grad.numpy()
        print(name, y.shape, float(y.abs().mean()))
    np.savez_compressed(os.path.join(HERE, "attention.npz"), **out)


if __name__ == "__main__":
    main()
