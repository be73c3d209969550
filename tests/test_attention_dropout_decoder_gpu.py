"""GPU: (1) attention-probability dropout fused into the tcgen05 kernels (utils.py:311,326): mask statistics, the forward
against `dropped probabilities @ v`, and the two backward kernels against autograd through the SAME mask; (2) DecoderLayer
(both `last` settings) and PositionwiseFeedForward against golden vectors made by the reference's own classes
(tests/golden/make_golden_decoder.py); (3) the reference-default constructors train."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_decoder import DEC_CASES, PFF_CASE, dec_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 2e-4


def _rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("BH,Lq,Lk,p,mask_kind", [(2, 256, 384, 0.1, None), (4, 300, 515, 0.25, "key"), (2, 129, 1000, 0.5, "full")])
def test_dropout_mask_forward_backward(cuda, BH, Lq, Lk, p, mask_kind):
    from isa_b200.attention import scaled_dot_product_attention
    d = dv = 12
    torch.manual_seed(BH + Lq)
    q = torch.randn(BH, Lq, d, dtype=torch.float64)
    k = torch.randn(BH, Lk, d, dtype=torch.float64)
    v = torch.randn(BH, Lk, dv, dtype=torch.float64)
    mask = None
    if mask_kind == "key":
        mask = torch.rand(2, 1, Lk) < 0.3
        mask[:, :, 0] = False
    elif mask_kind == "full":
        mask = torch.rand(2, Lq, Lk) < 0.3
        mask[:, :, 0] = False
    T = float(np.sqrt(d))
    mg = mask.to(cuda) if mask is not None else None
    qg, kg, vg = [t.float().to(cuda).requires_grad_(True) for t in (q, k, v)]
    seed = 1234567
    out, attn_d = scaled_dot_product_attention(qg, kg, vg, T, mg, return_attn=True, dropout_p=p, dropout_seed=seed)
    _, attn_0 = scaled_dot_product_attention(qg.detach(), kg.detach(), vg.detach(), T, mg, return_attn=True)
    a0, ad = attn_0.double().cpu(), attn_d.double().cpu()
    live = a0 > 0
    keep = (ad > 0) & live
    thr = round(p * 65536)
    scale = 65536.0 / (65536.0 - thr)
    # ---- the mask: keep rate, scaling of the survivors, rough independence along queries / keys, a new seed = a new mask
    n_live = int(live.sum())
    rate = float(keep.sum()) / n_live
    assert abs(rate - (1 - thr / 65536.0)) < 4.0 * np.sqrt(p * (1 - p) / n_live) + 1e-4, rate
    assert float((ad[keep] - a0[keep] * scale).abs().max()) < 1e-5
    assert float(ad[live & ~keep].abs().max()) == 0.0
    kf = keep.float()
    per_row = kf.sum(2) / live.float().sum(2).clamp_min(1)
    assert float(per_row.std()) < 3.0 * np.sqrt(p * (1 - p) / Lk) + 0.02      # no dead / always-kept rows
    if mask is None:
        a, b = kf[:, :, :-1].flatten(), kf[:, :, 1:].flatten()
        assert abs(float(((a - a.mean()) * (b - b.mean())).mean() / (a.std() * b.std()))) < 0.02   # neighbouring keys
        a, b = kf[:, :-1].flatten(), kf[:, 1:].flatten()
        assert abs(float(((a - a.mean()) * (b - b.mean())).mean() / (a.std() * b.std()))) < 0.02   # neighbouring queries
    _, attn_d2 = scaled_dot_product_attention(qg.detach(), kg.detach(), vg.detach(), T, mg, return_attn=True, dropout_p=p, dropout_seed=seed + 1)
    assert float(((attn_d2 > 0) != (attn_d > 0)).float().mean()) > 0.5 * p * (1 - p)
    # ---- forward = (dropped probabilities) @ v; backward = autograd through the same mask
    qr, kr, vr = [t.clone().requires_grad_(True) for t in (q, k, v)]
    s = torch.bmm(qr, kr.transpose(1, 2)) / T
    if mask is not None:
        s = s.masked_fill(mask.repeat(BH // 2, 1, 1), -np.inf)
    pr = torch.softmax(s, dim=2) * keep.double() * scale
    oref = torch.bmm(pr, vr)
    assert _rel(out, oref) < RTOL
    go = torch.randn_like(oref)
    oref.backward(go)
    out.backward(go.float().to(cuda))
    assert _rel(qg.grad, qr.grad) < RTOL
    assert _rel(kg.grad, kr.grad) < RTOL
    assert _rel(vg.grad, vr.grad) < RTOL


def test_reference_default_constructors_train(cuda):
    """MultiHeadAttention(n_head, d_model, d_k, d_v) / DecoderLayer(...) as the reference builds them (attn_dropout 0.1,
    dropout 0.1) run a training step; eval mode is deterministic and equals the dropout-free result."""
    from isa_b200.attention import DecoderLayer, MultiHeadAttention
    torch.manual_seed(0)
    mha = MultiHeadAttention(2, 24, 12, 12).to(cuda)
    x = torch.randn(2, 200, 24, device=cuda, requires_grad=True)
    mha.train()
    y1, _ = mha(x, x, x)
    y2, _ = mha(x, x, x)
    assert torch.isfinite(y1).all() and not torch.equal(y1, y2)       # fresh dropout masks per call
    y1.sum().backward()
    assert torch.isfinite(x.grad).all() and all(torch.isfinite(p.grad).all() for p in mha.parameters())
    mha.eval()
    e1, _ = mha(x, x, x)
    e2, _ = mha(x, x, x)
    assert torch.equal(e1, e2)
    layer = DecoderLayer(24, 40, 2, 12, 12).to(cuda).train()
    dec = torch.randn(2, 4, 24, device=cuda, requires_grad=True)
    enc = torch.randn(2, 150, 24, device=cuda)
    fg = torch.ones(2, 150, dtype=torch.uint8, device=cuda)
    out, _, _ = layer(dec, enc, fg)
    out.sum().backward()
    assert torch.isfinite(dec.grad).all()


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "decoder.npz"))


@pytest.mark.parametrize("case", DEC_CASES, ids=[c[0] for c in DEC_CASES])
def test_decoder_layer_golden(cuda, golden, case):
    from isa_b200.attention import DecoderLayer
    name, seed, b, Ld, Le, d_model, d_inner, n_head, d_k, d_v, last = case
    layer = DecoderLayer(d_model, d_inner, n_head, d_k, d_v, last=last).to(cuda).eval()
    sd = {k[len(name) + 3:]: torch.tensor(golden[k]) for k in golden.files if k.startswith(name + "_w_")}
    layer.load_state_dict(sd)
    dec, enc, fg = dec_inputs(case)
    td = torch.tensor(dec, device=cuda, requires_grad=True)
    te = torch.tensor(enc, device=cuda, requires_grad=True)
    y, _, _ = layer(td, te, torch.tensor(fg, device=cuda))
    assert tuple(y.shape) == golden[name + "_y"].shape
    assert _rel(y, golden[name + "_y"]) < RTOL
    y.backward(torch.tensor(golden[name + "_gy"], device=cuda))
    assert _rel(td.grad, golden[name + "_gdec"]) < 5e-4
    assert _rel(te.grad, golden[name + "_genc"]) < 5e-4
    # parameter gradients: relative to the largest gradient of the layer (the key-projection biases have an exactly zero
    # gradient -- softmax is shift invariant -- so the reference holds 1e-7 rounding noise there)
    gmax = max(float(np.abs(golden[k]).max()) for k in golden.files if k.startswith(name + "_g_"))
    for k_, p_ in layer.named_parameters():
        key = name + "_g_" + k_
        if key in golden.files:
            g = torch.tensor(golden[key]).double()
            got = p_.grad.double().cpu() if p_.grad is not None else torch.zeros_like(g)
            assert float((got - g).abs().max()) < 5e-4 * max(float(g.abs().max()), 1e-2 * gmax), k_


def test_positionwise_feed_forward_golden(cuda, golden):
    from isa_b200.attention import PositionwiseFeedForward
    name, seed, b, L, d_in, d_hid = PFF_CASE
    pff = PositionwiseFeedForward(d_in, d_hid).to(cuda).eval()
    pff.load_state_dict({k[len(name) + 3:]: torch.tensor(golden[k]) for k in golden.files if k.startswith(name + "_w_")})
    x = torch.tensor(golden[name + "_x"], device=cuda, requires_grad=True)
    y = pff(x)
    assert _rel(y, golden[name + "_y"]) < 1e-5
    y.backward(torch.tensor(golden[name + "_gy"], device=cuda))
    assert _rel(x.grad, golden[name + "_gx"]) < 1e-5
    for k_, p_ in pff.named_parameters():
        assert _rel(p_.grad, golden[name + "_g_" + k_]) < 1e-5, k_
