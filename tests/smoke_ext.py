"""Extended smoke check run by __graft_entry__.smoke() on cuda:0 (kept under tests/, outside the product package): one small invocation of EVERY kernel family of the hot
path, each asserted against its oracle (oracle/ is test infrastructure: it is imported here, in the checker, never by
the product modules).  The point is driver-visible proof: these launches are what GPUTEST.launches lists."""
import numpy as np
import torch


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(dev):
    from isa_b200 import clustering, synth
    from isa_b200.attention import MultiHeadAttention
    from isa_b200.renet import ReNet
    from isa_b200.seg_losses import SegLosses
    from oracle import kmeans as KM
    from oracle import seg_losses_ref
    from oracle.attention_ref import MHARef
    from oracle.renet_ref import ReNetRef

    # ---- ReNet: projection GEMMs (tcgen05) + GRU scans (tcgen05), forward and backward, against nn.GRU
    torch.manual_seed(1)
    ref = ReNetRef(32, 100, (1, 1)).double()
    mod = ReNet(32, 100, (1, 1)).to(dev)
    mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    x = torch.randn(2, 32, 12, 16, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    xg = x.float().to(dev).requires_grad_(True)
    yg = mod(xg)
    yg.backward(gy.float().to(dev))
    assert _rel(yg, yr) < 2e-4 and _rel(xg.grad, xr.grad) < 2e-4, "ReNet"
    assert _rel(mod.rnn_hor.weight_ih_l0.grad, ref.rnn_hor.weight_ih_l0.grad) < 2e-4, "ReNet weight gradient"
    print("smoke: ReNet fwd %.1e, dx %.1e" % (_rel(yg, yr), _rel(xg.grad, xr.grad)))

    # ---- multi-head attention (tcgen05 flash forward + two backward kernels, residual + LayerNorm kernel)
    torch.manual_seed(2)
    mref = MHARef(2, 24, 12, 12).double().eval()
    mha = MultiHeadAttention(2, 24, 12, 12, dropout=0.0, attn_dropout=0.0).to(dev)
    mha.load_state_dict({k: v.float() for k, v in mref.state_dict().items()})
    q = torch.randn(2, 300, 24, dtype=torch.float64)
    qr = q.clone().requires_grad_(True)
    orf, _ = mref(qr, qr, qr)
    go = torch.randn_like(orf)
    orf.backward(go)
    qg = q.float().to(dev).requires_grad_(True)
    og, _ = mha(qg, qg, qg)
    og.backward(go.float().to(dev))
    assert _rel(og, orf) < 2e-4 and _rel(qg.grad, qr.grad) < 5e-4, "attention"
    print("smoke: attention fwd %.1e, dq %.1e" % (_rel(og, orf), _rel(qg.grad, qr.grad)))

    # ---- fused CE + Dice
    rs = np.random.RandomState(3)
    logits = (2.0 * rs.standard_normal((2, 2, 24, 40))).astype(np.float32)
    cls = rs.randint(0, 2, size=(2, 24, 40)).astype(np.uint8)
    z = torch.tensor(logits, device=dev, requires_grad=True)
    ce, dice = SegLosses()(z, torch.tensor(cls, device=dev), time=1)
    (ce + dice).backward()
    o = seg_losses_ref.seg_losses(logits, np.stack([cls == 0, cls == 1], 1).astype(np.int64), time=1)
    assert abs(float(ce) - o["ce"]) < 1e-5 * o["ce"] and abs(float(dice) - o["dice"]) < 1e-5, "seg losses"
    assert np.abs(z.grad.cpu().numpy() - o["grad"]).max() < 1e-4 * np.abs(o["grad"]).max(), "seg losses gradient"

    # ---- inference epilogue: argmax + foreground compaction, k-means++ / Lloyd, label scatter + up-sampling
    d = synth.batch(5, 1, 24, 48, 64, 8, n_min=8, n_max=8, pull=0.8)
    lab = d["labels"][0]
    sem = np.stack([(lab == 255).astype(np.float32) * 0.8 + 0.1, (lab != 255).astype(np.float32) * 0.8 + 0.1])
    fg, X = KM.gather_foreground(sem, d["emb"][0])
    want = KM.scatter_labels(fg, KM.kmeans_oracle(X, 8, seed=0, n_init=6)["labels"])
    cls_map, ins_small, ins_up, cls_up, res = clustering.cluster_embeddings(
        torch.tensor(sem, device=dev), torch.tensor(d["emb"][0], device=dev), 8, 96, 100, seed=0, n_init=6)
    res.check()
    assert np.array_equal(ins_small.cpu().numpy(), want), "k-means labels differ from the oracle"
    assert np.array_equal(cls_map.cpu().numpy(), fg), "foreground map"
    assert np.array_equal(ins_up.cpu().numpy(), KM.upsample_nearest(want, 96, 100)), "up-sampled mask"
    print("smoke: clustering bit-exact against the oracle (%d points, %d Lloyd iterations)" % (len(X), int(res.n_iter.sum())))
