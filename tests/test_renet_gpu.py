"""GPU: ReNet (persistent GRU scan kernels through the C-ABI) against the nn.GRU oracle
(oracle/renet_ref.py) on identical weights: forward, input gradient and every parameter gradient.
Tolerance: 1e-3 relative (north_star) -- asserted at 2e-4."""
import numpy as np
import pytest
import torch

from oracle.renet_ref import ReNetRef

pytestmark = pytest.mark.gpu

RTOL = 2e-4


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("B,C,H,W,n,patch", [(2, 16, 8, 12, 8, (1, 1)), (1, 3, 16, 16, 16, (2, 2)), (3, 32, 20, 9, 100, (1, 1)),
                                             (2, 7, 5, 33, 12, (1, 1)), (16, 64, 16, 16, 32, (1, 1)), (1, 5, 7, 9, 4, (2, 3))])
def test_renet_forward_backward(cuda, B, C, H, W, n, patch):
    from isa_b200.renet import ReNet
    torch.manual_seed(B * 100 + n)
    ref = ReNetRef(C, n, patch).double()
    mod = ReNet(C, n, patch).to(cuda)
    mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    x = torch.randn(B, C, H, W, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    xg = x.float().to(cuda).requires_grad_(True)
    yg = mod(xg)
    assert yg.shape == yr.shape
    assert _rel(yg, yr) < RTOL
    yg.backward(gy.float().to(cuda))
    assert _rel(xg.grad, xr.grad) < RTOL
    for (name, p), (_, q) in zip(mod.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad, q.grad) < RTOL, name


def test_renet_inference_no_stash(cuda):
    from isa_b200.renet import ReNet
    torch.manual_seed(0)
    ref = ReNetRef(24, 20)
    mod = ReNet(24, 20).to(cuda)
    mod.load_state_dict(ref.state_dict())
    x = torch.randn(2, 24, 10, 14)
    with torch.no_grad():
        assert _rel(mod(x.to(cuda)), ref(x)) < RTOL


def test_state_dict_names_match_nn_gru():
    from isa_b200.renet import ReNet
    names = set(ReNet(8, 4).state_dict().keys())
    assert {"rnn_hor.weight_ih_l0", "rnn_hor.weight_hh_l0_reverse", "rnn_ver.bias_hh_l0", "rnn_ver.weight_ih_l0_reverse"} <= names
