"""Golden vectors of DecoderLayer (both `last` settings) and PositionwiseFeedForward from the REFERENCE CLASSES themselves
(/root/reference/code/lib/archs/modules/utils.py:138-164, 229-246, imported through oracle/ref_loader.py, eval mode so
that nn.Dropout is the identity): outputs, input gradients and every parameter gradient.  Build container only:
    python tests/golden/make_golden_decoder.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
from oracle import ref_loader  # noqa: E402

# name, seed, b, L_dec, L_enc, d_model, d_inner, n_head, d_k, d_v, last
DEC_CASES = [
    ("dec", 0, 2, 5, 300, 24, 40, 2, 12, 12, False),
    ("dec_self", 1, 1, 130, 130, 24, 40, 2, 12, 12, False),
    ("dec_last", 2, 2, 3, 200, 24, 40, 2, 12, 12, True),
]
PFF_CASE = ("pff", 3, 3, 77, 24, 40)


def dec_inputs(case):
    name, seed, b, Ld, Le, d_model, d_inner, n_head, d_k, d_v, last = case
    rs = np.random.RandomState(seed)
    dec = rs.standard_normal((b, Ld, d_model)).astype(np.float32)
    enc = rs.standard_normal((b, Le, d_model)).astype(np.float32)
    fg = (rs.uniform(size=(b, Le)) < 0.6).astype(np.uint8)      # 1 = foreground = attended (utils.py:151-152)
    fg[:, 0] = 1
    return dec, enc, fg


def main():
    U = ref_loader.attention_utils()
    out = {}
    for case in DEC_CASES:
        name, seed, b, Ld, Le, d_model, d_inner, n_head, d_k, d_v, last = case
        torch.manual_seed(seed)
        layer = U.DecoderLayer(d_model, d_inner, n_head, d_k, d_v, last=last).eval()
        dec, enc, fg = dec_inputs(case)
        td, te = torch.tensor(dec, requires_grad=True), torch.tensor(enc, requires_grad=True)
        y, a_slf, a_enc = layer(td, te, torch.tensor(fg))
        gy = torch.tensor(np.random.RandomState(100 + seed).standard_normal(tuple(y.shape)).astype(np.float32))
        y.backward(gy)
        out[name + "_y"] = y.detach().numpy()
        out[name + "_gy"] = gy.numpy()
        out[name + "_gdec"] = td.grad.numpy()
        out[name + "_genc"] = te.grad.numpy()
        for k_, v_ in layer.state_dict().items():
            out[name + "_w_" + k_] = v_.numpy()
        for k_, p_ in layer.named_parameters():
            if p_.grad is not None:
                out[name + "_g_" + k_] = p_.grad.numpy()
        print(name, tuple(y.shape), float(y.abs().mean()))
    name, seed, b, L, d_in, d_hid = PFF_CASE
    torch.manual_seed(seed)
    pff = U.PositionwiseFeedForward(d_in, d_hid).eval()
    x = np.random.RandomState(seed).standard_normal((b, L, d_in)).astype(np.float32)
    tx = torch.tensor(x, requires_grad=True)
    y = pff(tx)
    gy = torch.tensor(np.random.RandomState(100 + seed).standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)
    out[name + "_x"] = x
    out[name + "_y"] = y.detach().numpy()
    out[name + "_gy"] = gy.numpy()
    out[name + "_gx"] = tx.grad.numpy()
    for k_, v_ in pff.state_dict().items():
        out[name + "_w_" + k_] = v_.numpy()
    for k_, p_ in pff.named_parameters():
        out[name + "_g_" + k_] = p_.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "decoder.npz"), **out)


if __name__ == "__main__":
    main()
