"""GPU: spatial / channel / local attention layers (sm_100a kernels through the C-ABI, reference-named modules)
against golden vectors from the reference's own classes (tests/golden/spatial.npz: outputs, input gradients
and every parameter gradient, training mode) and against oracle/spatial_ref.py in fp64 on random shapes,
including empty masks, H*W not divisible by 4, uint8 / float32 / bool masks, and one large map.
Tolerance: north_star 1e-3 relative for fp32; asserted at 1e-4."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import spatial_ref as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_spatial import CASES, D_H, D_K, D_V, N_HEAD, inputs  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(autouse=True)
def _fp32_library_convs():
    """The 1x1 / 3x3 convolutions around the kernels are library (cuDNN) calls; keep them in true fp32 so the
    comparison measures the kernels, not cuDNN's TF32 default."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "spatial.npz"))


def _rel(a, b):
    a = a.detach().double().cpu() if torch.is_tensor(a) else torch.tensor(np.asarray(a)).double()
    b = b.detach().double().cpu() if torch.is_tensor(b) else torch.tensor(np.asarray(b)).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _load(mod, G, prefix):
    sd = {k: torch.tensor(G[prefix + "_w_" + k]) for k in mod.state_dict().keys()}
    mod.load_state_dict(sd)
    return mod


def _check_param_grads(mod, G, prefix, tol=RTOL, atol=1e-4):
    """Parameter gradients are sums over every pixel of O(1) terms; some are mathematically zero (a bias in front
    of a softmax or an instance norm) or the remainder of a large cancellation (a bias in front of maskBN), so
    they are compared with an absolute floor `atol` (the fp32 summation noise of either implementation)."""
    for k, p in mod.named_parameters():
        key = "%s_g_%s" % (prefix, k)
        if key in G.files:
            assert p.grad is not None, k
            ref = torch.tensor(G[key]).double()
            err = float((p.grad.detach().double().cpu() - ref).abs().max())
            assert err <= tol * float(ref.abs().max()) + atol, (k, err, float(ref.abs().max()))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("tag,red", [("sp", D_H), ("sp2", 2)])
def test_spatial_attention_layer_golden(cuda, G, case, tag, red):
    from isa_b200.spatial_attention import SpatialAttentionLayer
    name, seed, b, c, h, w, n, n_max = case
    base, sem, ins, q = inputs(case)
    pre = "%s_%s" % (name, tag)
    m = _load(SpatialAttentionLayer(c, red), G, pre).to(cuda).train()
    B = torch.tensor(base, device=cuda, requires_grad=True)
    y = m(B, torch.tensor(sem, device=cuda))
    assert _rel(y, G[pre + "_y"]) < RTOL
    y.backward(torch.tensor(G[pre + "_gy"], device=cuda))
    assert _rel(B.grad, G[pre + "_gbase"]) < RTOL
    _check_param_grads(m, G, pre)
    m2 = _load(SpatialAttentionLayer(c, red, multiply=False), G, pre).to(cuda).train()
    assert _rel(m2(torch.tensor(base, device=cuda), torch.tensor(sem, device=cuda)), G[pre + "_beta"]) < RTOL


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_hard_attention_layer_golden(cuda, G, case):
    from isa_b200.spatial_attention import HardAttentionLayer
    name, seed, b, c, h, w, n, n_max = case
    base, sem, ins, q = inputs(case)
    pre = name + "_hard"
    m = _load(HardAttentionLayer(c, D_K, D_H), G, pre).to(cuda).train()
    S = torch.tensor(base, device=cuda, requires_grad=True)
    e_split, e_org = m(S, torch.tensor(sem, device=cuda), torch.tensor(ins, device=cuda))
    assert _rel(e_split, G[pre + "_split"]) < RTOL
    assert _rel(e_org, G[pre + "_org"]) < RTOL
    assert float(e_split[:, n_max:].abs().max()) == 0.0        # empty instance channels: NaN -> 0
    ((e_split * torch.tensor(G[pre + "_g1"], device=cuda)).sum() + (e_org * torch.tensor(G[pre + "_g2"], device=cuda)).sum()).backward()
    assert _rel(S.grad, G[pre + "_gS"]) < RTOL
    _check_param_grads(m, G, pre)
    # uint8 instance masks give the same result as the float one-hot
    with torch.no_grad():
        e2, _ = m.eval()(torch.tensor(base, device=cuda), torch.tensor(sem, device=cuda), torch.tensor(ins.astype(np.uint8), device=cuda))
        e3, _ = m(torch.tensor(base, device=cuda), torch.tensor(sem, device=cuda), torch.tensor(ins, device=cuda))
    assert torch.equal(e2, e3)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_se_layer_and_readout_golden(cuda, G, case):
    from isa_b200.spatial_attention import AttentionLayer, Decoder
    name, seed, b, c, h, w, n, n_max = case
    base, sem, ins, q = inputs(case)
    m = _load(AttentionLayer(c), G, name + "_se").to(cuda).train()
    X = torch.tensor(base, device=cuda, requires_grad=True)
    y = m(X)
    assert _rel(y, G[name + "_se_y"]) < RTOL
    y.backward(torch.tensor(G[name + "_se_gy"], device=cuda))
    assert _rel(X.grad, G[name + "_se_gx"]) < RTOL
    _check_param_grads(m, G, name + "_se")
    dec = Decoder(1, c, 40, N_HEAD, D_K, D_V).to(cuda)
    Q = torch.tensor(q, device=cuda, requires_grad=True)
    E = torch.tensor(base, device=cuda, requires_grad=True)
    out = dec(Q, E, None)
    assert _rel(out, G[name + "_ro_y"]) < RTOL
    out.backward(torch.tensor(G[name + "_ro_gy"], device=cuda))
    assert _rel(Q.grad, G[name + "_ro_gq"]) < RTOL
    assert _rel(E.grad, G[name + "_ro_genc"]) < RTOL


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("dil", [1, 3])
def test_local_attention_block_golden(cuda, G, case, dil):
    from isa_b200.spatial_attention import _AttenAsppBlock
    name, seed, b, c, h, w, n, n_max = case
    base, sem, ins, q = inputs(case)
    pre = "%s_la%d" % (name, dil)
    m = _load(_AttenAsppBlock(dil, c, D_K, D_V, 40, N_HEAD), G, pre).to(cuda).train()
    X = torch.tensor(base, device=cuda, requires_grad=True)
    y = m(X, torch.tensor(sem, device=cuda))
    assert _rel(y, G[pre + "_y"]) < 2e-4
    y.backward(torch.tensor(G[pre + "_gy"], device=cuda))
    assert _rel(X.grad, G[pre + "_gx"]) < 5e-4      # two instance norms amplify fp32 rounding of either implementation
    _check_param_grads(m, G, pre, 5e-4)


@pytest.mark.parametrize("B,K,HW,mkind,scaled,n2z", [
    (1, 1, 4096, "f32", True, False), (2, 5, 999, "u8", False, True), (3, 32, 65536, "f32", False, True),
    (2, 3, 8192 * 3 + 4, "bool", True, True), (1, 2, 7, "u8", True, False), (16, 32, 4096, "u8", False, True)])
def test_masked_softmax_vs_oracle(cuda, B, K, HW, mkind, scaled, n2z):
    from isa_b200.spatial_attention import masked_softmax_hw
    torch.manual_seed(B * 100 + K)
    x = torch.randn(B, HW, dtype=torch.float64) * 3
    mask = (torch.rand(B, K, HW) < 0.4)
    if n2z:
        mask[0, K - 1] = False                      # an empty instance
    mask[0, 0, 0] = True
    scale = (mask.sum(2).double() if scaled else None)
    xr = x.clone().requires_grad_(True)
    yr = O.masked_softmax_hw_ref(xr, mask.double(), scale, n2z)
    gy = torch.randn_like(yr)
    (yr * gy).sum().backward()
    mg = {"f32": mask.float(), "u8": mask.to(torch.uint8), "bool": mask}[mkind].to(cuda)
    xg = x.float().to(cuda).requires_grad_(True)
    yg = masked_softmax_hw(xg, mg, scale.float().to(cuda) if scaled else None, n2z)
    assert _rel(yg, yr) < RTOL
    assert float(yg[~mask.to(cuda)].abs().max()) == 0.0
    (yg * gy.float().to(cuda)).sum().backward()
    assert _rel(xg.grad, xr.grad) < RTOL


def test_masked_softmax_empty_row_is_nan_like_the_reference(cuda):
    from isa_b200.spatial_attention import masked_softmax_hw
    x = torch.randn(2, 64, device=cuda)
    mask = torch.ones(2, 1, 64, device=cuda)
    mask[1] = 0
    y = masked_softmax_hw(x, mask, mask.sum(2), False)
    ref = O.masked_softmax_hw_ref(x.cpu(), mask.cpu(), mask.sum(2).cpu(), False)
    assert torch.isnan(y[1]).all() and torch.isnan(ref[1]).all()
    assert _rel(y[0], ref[0]) < RTOL


@pytest.mark.parametrize("Bh,dk,dv,h,w,dil,masked", [(2, 12, 12, 16, 16, 1, True), (4, 12, 12, 33, 21, 2, True), (1, 5, 7, 9, 40, 4, False),
                                                     (6, 20, 12, 64, 64, 3, True)])
def test_local_attention_vs_oracle(cuda, Bh, dk, dv, h, w, dil, masked):
    from isa_b200.spatial_attention import local_attention
    torch.manual_seed(Bh * 10 + dil)
    Q, K = torch.randn(Bh, dk, h, w, dtype=torch.float64), torch.randn(Bh, dk, h, w, dtype=torch.float64)
    V = torch.randn(Bh, dv, h, w, dtype=torch.float64)
    mb = 2 if Bh % 2 == 0 else 1
    nomask = (torch.rand(mb, 1, h, w) < 0.5).double() if masked else None
    if masked:
        nomask[0, 0, 2:7, 2:9] = 1     # a fully excluded neighbourhood (softmax NaN -> 0)
    Qr, Kr, Vr = [t.clone().requires_grad_(True) for t in (Q, K, V)]
    ref = O.local_attention_ref(Qr, Kr, Vr, nomask.repeat(Bh // mb, 1, 1, 1) if masked else None, dil, dk ** -0.5)
    g = torch.randn_like(ref)
    ref.backward(g)
    Qg, Kg, Vg = [t.float().to(cuda).requires_grad_(True) for t in (Q, K, V)]
    out = local_attention(Qg, Kg, Vg, nomask.float().to(cuda) if masked else None, dil, dk ** -0.5)
    assert _rel(out, ref) < RTOL
    out.backward(g.float().to(cuda))
    assert _rel(Qg.grad, Qr.grad) < RTOL
    assert _rel(Kg.grad, Kr.grad) < RTOL
    assert _rel(Vg.grad, Vr.grad) < RTOL


@pytest.mark.parametrize("b,c,h,w", [(2, 32, 64, 64), (1, 6, 5, 7), (16, 32, 256, 256)])
def test_squeeze_excite_vs_oracle(cuda, b, c, h, w):
    from isa_b200.spatial_attention import AttentionLayer
    torch.manual_seed(b + c)
    m = AttentionLayer(c).to(cuda)
    x = torch.randn(b, c, h, w)
    p = [t.detach().double().cpu().requires_grad_(True) for t in (m.fc[0].weight, m.fc[0].bias, m.fc[2].weight, m.fc[2].bias)]
    xr = x.double().requires_grad_(True)
    yr = O.squeeze_excite_ref(xr, *p)
    g = torch.randn_like(yr)
    yr.backward(g)
    xg = x.to(cuda).requires_grad_(True)
    yg = m(xg)
    assert _rel(yg, yr) < RTOL
    yg.backward(g.float().to(cuda))
    assert _rel(xg.grad, xr.grad) < RTOL
    for q_, r_ in zip((m.fc[0].weight, m.fc[0].bias, m.fc[2].weight, m.fc[2].bias), p):
        assert _rel(q_.grad, r_.grad) < RTOL


@pytest.mark.parametrize("b,c,h,w,per_channel_mask", [(4, 1, 64, 64, False), (3, 5, 33, 47, False), (2, 4, 40, 24, True), (16, 1, 256, 256, False)])
def test_mask_bn_vs_oracle(cuda, b, c, h, w, per_channel_mask):
    """maskBN's training branch (utils.py:573-586) as kernels: output, batch statistics / running averages, and the gradient
    THROUGH the masked mean and variance (input, weight, bias) against the float64 autograd of the restated formula."""
    from isa_b200.spatial_attention import maskBN
    torch.manual_seed(7 * b + c)
    m = maskBN(c).to(cuda).train()
    x = torch.randn(b, c, h, w) * 1.7 + 0.6
    mask = (torch.rand(b, c if per_channel_mask else 1, h, w) > 0.4).float()
    mask[0] = 0                                                     # an image without foreground: mm = 1
    wr = m.weight.detach().double().cpu().requires_grad_(True)
    br = m.bias.detach().double().cpu().requires_grad_(True)
    xr = x.double().requires_grad_(True)
    yr = O.mask_bn_train_ref(xr, mask.double(), wr, br, m.eps)
    g = torch.randn_like(yr)
    yr.backward(g)
    xg = x.to(cuda).requires_grad_(True)
    yg = m(xg, mask.to(cuda))
    assert _rel(yg, yr) < RTOL
    yg.backward(g.float().to(cuda))
    assert _rel(xg.grad, xr.grad) < RTOL
    assert _rel(m.weight.grad, wr.grad) < RTOL and _rel(m.bias.grad, br.grad) < RTOL
    # running averages: the reference's convention running * momentum + (1 - momentum) * batch (utils.py:580-581)
    mm = mask.double().expand(b, c, h, w).reshape(b, -1).sum(1) + 1
    x2, m2 = x.double().reshape(b, c, -1), mask.double().expand(b, c, h, w).reshape(b, c, -1)
    mean = ((x2 * m2).sum(2) / mm[:, None]).mean(0)
    var = ((((x2 - mean[None, :, None]) ** 2) * m2).sum(2) / mm[:, None]).mean(0)
    assert _rel(m.running_mean, 0.1 * torch.zeros(c, dtype=torch.float64) + 0.9 * mean) < RTOL
    assert _rel(m.running_var, 0.1 * torch.ones(c, dtype=torch.float64) + 0.9 * var) < RTOL
    # deterministic: a second call gives bit-identical results
    m2_ = maskBN(c).to(cuda).train()
    m2_.load_state_dict({k_: v_ for k_, v_ in m.state_dict().items()})
    assert torch.equal(m2_(x.to(cuda), mask.to(cuda)).detach(), m2_(x.to(cuda), mask.to(cuda)).detach())
