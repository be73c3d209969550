"""CPU: the attention restatement (oracle/attention_ref.py) against golden vectors produced by the
reference's own MultiHeadAttention class (tests/golden/make_golden_attn.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle.attention_ref import MHARef

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_attn import ATTN_CASES, case_inputs  # noqa: E402


def load_weights(mod, g, name):
    sd = {k: torch.tensor(g[name + "_w_" + k]) for k in mod.state_dict().keys()}
    mod.load_state_dict(sd)


@pytest.mark.parametrize("case", ATTN_CASES, ids=[c[0] for c in ATTN_CASES])
def test_mha_oracle_matches_reference_golden(golden_dir, case):
    name, seed, b, Lq, Lk, n_head, d_model, d_k, d_v, masked = case
    g = np.load(os.path.join(golden_dir, "attention.npz"))
    mod = MHARef(n_head, d_model, d_k, d_v)
    load_weights(mod, g, name)
    q, kv, mask = case_inputs(case)
    tq = torch.tensor(q, requires_grad=True)
    tkv = tq if kv is q else torch.tensor(kv)
    y, attn = mod(tq, tkv, tkv, torch.tensor(mask) if mask is not None else None)
    y.backward(torch.tensor(g[name + "_gy"]))
    np.testing.assert_allclose(y.detach().numpy(), g[name + "_y"], atol=2e-5)
    np.testing.assert_allclose(attn.detach().sum(2).numpy(), g[name + "_attn_rowsum"], atol=1e-5)
    np.testing.assert_allclose(tq.grad.numpy(), g[name + "_gq"], atol=2e-5)
