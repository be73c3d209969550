"""Golden vectors of the spatial / channel / local attention layers from the REFERENCE CLASSES themselves
(lib/archs/modules/utils.py: SpatialAttentionLayer, HardAttentionLayer (+maskBN), AttentionLayer, Decoder,
_ScalePDAttention, make_position_encoding; lib/archs/reseg.py ReSeg.set_position_encoding), imported from
/root/reference through oracle/ref_loader.py, training mode (batch statistics), CPU fp32.
Run through `python tests/golden/make_golden.py spatial`.  Inputs are regenerated from seeds by `inputs()`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
from isa_b200 import synth  # noqa: E402

# name, seed, b, c, h, w, n (instance channels), n_max (objects present; channels beyond are empty masks)
CASES = [("a", 0, 2, 24, 20, 24, 8, 5), ("b", 1, 3, 24, 17, 13, 32, 12)]
D_K, D_H, N_HEAD, D_V = 12, 20, 2, 12


def inputs(case):
    name, seed, b, c, h, w, n, n_max = case
    rs = np.random.RandomState(1000 + seed)
    d = synth.batch(seed, b, c, h, w, n, n_min=2, n_max=n_max)
    base = d["emb"]
    sem = (d["labels"] != 255).astype(np.float32)[:, None]
    ins = synth.onehot(d["labels"], n)
    q = rs.standard_normal((b, c)).astype(np.float32)
    return base, sem, ins, q


def _sd(prefix, mod, out):
    for k, v in mod.state_dict().items():
        out[prefix + "_w_" + k] = v.detach().numpy().copy()


def main():
    from oracle import ref_loader
    U = ref_loader.attention_utils()
    out = {}
    for case in CASES:
        name, seed, b, c, h, w, n, n_max = case
        base, sem, ins, q = inputs(case)
        rs = np.random.RandomState(2000 + seed)

        # --- A6 SpatialAttentionLayer(d_model, d_h) as the live decoder builds it (attenet2.py:29): reduction = d_h
        for tag, red in (("sp", D_H), ("sp2", 2)):
            torch.manual_seed(seed)
            m = U.SpatialAttentionLayer(c, red).train()
            _sd(name + "_" + tag, m, out)
            cap = {}
            hk = m.spatial_fc.register_forward_hook(lambda mod, i, o: cap.__setitem__("beta", o.detach().clone()))
            B = torch.tensor(base, requires_grad=True)
            y = m(B, torch.tensor(sem))
            hk.remove()
            gy = torch.tensor(rs.standard_normal(y.shape).astype(np.float32))
            y.backward(gy)
            out["%s_%s_y" % (name, tag)] = y.detach().numpy()
            out["%s_%s_gy" % (name, tag)] = gy.numpy()
            out["%s_%s_gbase" % (name, tag)] = B.grad.numpy()
            out["%s_%s_beta_logits" % (name, tag)] = cap["beta"].numpy()
            for k_, p_ in m.named_parameters():
                out["%s_%s_g_%s" % (name, tag, k_)] = p_.grad.numpy()
            m2 = U.SpatialAttentionLayer(c, red, multiply=False).train()
            m2.load_state_dict(m.state_dict())
            out["%s_%s_beta" % (name, tag)] = m2(torch.tensor(base), torch.tensor(sem)).detach().numpy()

        # --- A7 HardAttentionLayer (+ maskBN)
        torch.manual_seed(seed + 10)
        m = U.HardAttentionLayer(c, D_K, D_H).train()
        _sd(name + "_hard", m, out)
        S = torch.tensor(base, requires_grad=True)
        e_split, e_org = m(S, torch.tensor(sem), torch.tensor(ins))
        g1 = torch.tensor(rs.standard_normal(e_split.shape).astype(np.float32))
        g2 = torch.tensor(rs.standard_normal(e_org.shape).astype(np.float32))
        ((e_split * g1).sum() + (e_org * g2).sum()).backward()
        out[name + "_hard_split"] = e_split.detach().numpy()
        out[name + "_hard_org"] = e_org.detach().numpy()
        out[name + "_hard_g1"] = g1.numpy()
        out[name + "_hard_g2"] = g2.numpy()
        out[name + "_hard_gS"] = S.grad.numpy()
        for k_, p_ in m.named_parameters():
            if p_.grad is not None:
                out["%s_hard_g_%s" % (name, k_)] = p_.grad.numpy()

        # --- A8 AttentionLayer (SE)
        torch.manual_seed(seed + 20)
        m = U.AttentionLayer(c).train()
        _sd(name + "_se", m, out)
        X = torch.tensor(base, requires_grad=True)
        y = m(X)
        gy = torch.tensor(rs.standard_normal(y.shape).astype(np.float32))
        y.backward(gy)
        out[name + "_se_y"] = y.detach().numpy()
        out[name + "_se_gy"] = gy.numpy()
        out[name + "_se_gx"] = X.grad.numpy()
        for k_, p_ in m.named_parameters():
            out["%s_se_g_%s" % (name, k_)] = p_.grad.numpy()

        # --- A4 Decoder readout
        torch.manual_seed(seed + 30)
        m = U.Decoder(1, c, 40, N_HEAD, D_K, D_V)
        Q = torch.tensor(q, requires_grad=True)
        E = torch.tensor(base, requires_grad=True)
        y = m(Q, E, None)
        gy = torch.tensor(rs.standard_normal(y.shape).astype(np.float32))
        y.backward(gy)
        out[name + "_ro_y"] = y.detach().numpy()
        out[name + "_ro_gy"] = gy.numpy()
        out[name + "_ro_gq"] = Q.grad.numpy()
        out[name + "_ro_genc"] = E.grad.numpy()

        # --- A5 _ScalePDAttention / _AttenAsppBlock (dilations 1 and 3)
        for dil in (1, 3):
            torch.manual_seed(seed + 40 + dil)
            m = U._AttenAsppBlock(dil, c, D_K, D_V, 40, N_HEAD).train()
            tag = "%s_la%d" % (name, dil)
            _sd(tag, m, out)
            X = torch.tensor(base, requires_grad=True)
            y = m(X, torch.tensor(sem))
            gy = torch.tensor(rs.standard_normal(y.shape).astype(np.float32))
            y.backward(gy)
            out[tag + "_y"] = y.detach().numpy()
            out[tag + "_gy"] = gy.numpy()
            out[tag + "_gx"] = X.grad.numpy()
            for k_, p_ in m.named_parameters():
                out["%s_g_%s" % (tag, k_)] = p_.grad.numpy()
        print(name, "done")

    # --- A9 position encodings (utils.make_position_encoding; ReSeg.set_position_encoding arithmetic, reseg.py:132-137)
    out["pe_1d"] = U.make_position_encoding(np, 2, 37, 24)
    hh, ww, nu = 9, 14, 24
    h_vec = np.tile(U.make_position_encoding(np, 1, hh, nu // 2, f=10000.)[:, :, :, np.newaxis], (1, 1, 1, ww))
    w_vec = np.tile(U.make_position_encoding(np, 1, ww, nu // 2, f=10000.)[:, :, np.newaxis, :], (1, 1, hh, 1))
    out["pe_2d"] = np.concatenate([h_vec, w_vec], axis=1)
    np.savez_compressed(os.path.join(HERE, "spatial.npz"), **out)
    print("spatial.npz: %d arrays, %.1f KB" % (len(out), os.path.getsize(os.path.join(HERE, "spatial.npz")) / 1e3))


if __name__ == "__main__":
    main()
