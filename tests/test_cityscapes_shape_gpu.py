"""GPU: BASELINE.json configs[3] -- Cityscapes-shaped 1024x2048 inputs, up to 64 instances, embedding width 32 --
through every hot-path kernel at FULL size.  Where the oracle finishes in seconds it is compared directly (loss,
k-means with a bounded restart / iteration budget); elsewhere size-independent properties are used (query-subset
equality for attention, oracle rows for the ReNet sweep, idempotence and partition identities for clustering)."""
import numpy as np
import pytest
import torch

from isa_b200 import synth
from oracle import disc_loss as O
from oracle import kmeans as KM

pytestmark = pytest.mark.gpu

H, W, K, C = 1024, 2048, 64, 32


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def scene():
    return synth.batch(2024, 1, C, H, W, K, n_min=48, n_max=64, fg_frac=0.3, pull=0.75)


def _loss_ref_label_map(emb, labels, n_obj):
    """The shipped composite (discriminative.py:162-188: 1.0 * variance term + 0.005 * q-regulariser, L2-normalised means)
    for ONE image with a hard label map, in O(P C) memory (the dense-mask oracle needs P*K*C doubles: 34 GB at this
    size).  float64 torch on the CPU with autograd; checked against oracle/disc_loss.py on a crop below."""
    Cn = emb.shape[0]
    x = torch.tensor(emb.reshape(Cn, -1).T.copy(), dtype=torch.float64, requires_grad=True)   # (P, C)
    lab = torch.tensor(labels.reshape(-1).astype(np.int64))
    fgm = lab != 255
    idx = lab[fgm]
    xs = x[fgm]
    sums = torch.zeros(n_obj, Cn, dtype=torch.float64).index_add_(0, idx, xs)
    cnt = torch.bincount(idx, minlength=n_obj).double()
    mu = sums / cnt[:, None]
    mu = mu / mu.norm(dim=1, keepdim=True)
    dist = (xs - mu[idx]).norm(dim=1)
    var = (torch.clamp(dist - 0.5, min=0.0) ** 2).sum() / cnt.sum()
    l = torch.zeros(x.shape[0], dtype=torch.float64)
    l = l.masked_scatter(fgm, xs.norm(dim=1))
    qreg = ((l - 1.0) ** 2).sum() / int(fgm.sum())
    loss = var + 0.005 * qreg
    loss.backward()
    return float(loss.detach()), mu.detach().numpy(), x.grad.T.reshape(emb.shape).numpy()


def test_discriminative_loss_full_size(cuda, scene):
    from isa_b200.losses import DiscriminativeLoss
    d = scene
    n_obj = int(d["n_objects"][0])
    # the slim reference restates the oracle: same numbers on a crop where the dense-mask oracle fits
    crop = (slice(0, 1), slice(None), slice(300, 428), slice(600, 856))
    lab_c = d["labels"][0, 300:428, 600:856]
    present = np.unique(lab_c[lab_c != 255])
    remap = np.full(256, 255, dtype=np.uint8)
    remap[present] = np.arange(len(present), dtype=np.uint8)
    lab_c = remap[lab_c]
    o = O.discriminative_loss(d["emb"][crop], lab_c[None], np.array([len(present)]), K, 0.5, 1.5, 2, want_grad=True)
    l_c, mu_c, g_c = _loss_ref_label_map(d["emb"][0][:, 300:428, 600:856], lab_c, len(present))
    assert abs(l_c - float(o["loss"])) < 1e-9 * abs(float(o["loss"]))
    assert _rel(g_c, o["grad"][0]) < 1e-9
    # full size
    l_ref, mu_ref, g_ref = _loss_ref_label_map(d["emb"][0], d["labels"][0], n_obj)
    for kind in ("label", "f32"):
        tgt = d["labels"] if kind == "label" else synth.onehot(d["labels"], K)
        x = torch.tensor(d["emb"], device=cuda, requires_grad=True)
        loss, means = DiscriminativeLoss(0.5, 1.5, 2)(x, torch.tensor(tgt, device=cuda), torch.tensor(d["n_objects"], device=cuda), K)
        loss.backward()
        assert abs(float(loss) - l_ref) <= 1e-4 * abs(l_ref)
        np.testing.assert_allclose(means.detach().cpu().numpy()[0, :n_obj], mu_ref, atol=1e-5)
        assert _rel(x.grad[0], g_ref) < 1e-4


def test_kmeans_full_size_bit_exact_bounded_budget(cuda, scene):
    """n_fg ~ 6e5 points, k = 64, C = 32: bit-exact against the oracle with a restart / iteration budget the CPU
    restatement finishes in seconds (the full 35 x 500 budget is covered through properties below)."""
    from isa_b200 import clustering
    lab = scene["labels"][0]
    sem = np.stack([(lab == 255).astype(np.float32), (lab != 255).astype(np.float32)])
    fg, X = KM.gather_foreground(sem, scene["emb"][0])
    assert 4e5 < len(X) < 9e5
    o = KM.kmeans_oracle(X, K, seed=5, n_init=2, max_iter=6)
    labels, res = clustering.kmeans_fit_predict(torch.tensor(X, device=cuda), K, seed=5, n_init=2, max_iter=6)
    assert np.array_equal(res.seed_idx.cpu().numpy(), o["seed_idx"])
    assert np.array_equal(res.n_iter.cpu().numpy(), o["n_iter"])
    assert np.array_equal(res.inertia.cpu().numpy(), o["inertia"])
    assert np.array_equal(labels.cpu().numpy(), o["labels"])
    assert np.array_equal(res.centers.cpu().numpy(), o["centers"])


def test_cluster_pipeline_full_size_properties(cuda, scene):
    """Full budget (n_init = 35, max_iter = 500) on the device: deterministic / idempotent, every foreground pixel gets
    a label in 1..k and every background pixel 0, the partition recovers the planted instances, and the label image
    survives the INTER_NEAREST round trip to the same size."""
    from isa_b200 import clustering
    lab = scene["labels"][0]
    n_obj = int(scene["n_objects"][0])
    sem = torch.tensor(np.stack([(lab == 255), (lab != 255)]).astype(np.float32), device=cuda)
    emb = torch.tensor(scene["emb"][0], device=cuda)
    out1 = clustering.cluster_embeddings(sem, emb, n_obj, H, W, seed=0)
    out2 = clustering.cluster_embeddings(sem, emb, n_obj, H, W, seed=0)
    out1[4].check()
    ins = out1[1].cpu().numpy()
    assert torch.equal(out1[1], out2[1]) and torch.equal(out1[2], out2[2])
    assert np.array_equal(ins != 0, lab != 255)
    assert ins.max() <= n_obj and ins[lab != 255].min() >= 1
    assert np.array_equal(out1[2].cpu().numpy(), ins)              # same-size up-sampling is the identity
    # objective check: the best of 35 restarts is at least as good (within 2 %) as the planted partition the embeddings
    # were drawn around, and every one of the k clusters is used
    X = scene["emb"][0][:, lab != 255].T.astype(np.float64)

    def inertia(labels):
        tot = 0.0
        for c in np.unique(labels):
            xc = X[labels == c]
            tot += ((xc - xc.mean(0)) ** 2).sum()
        return tot
    found, planted = ins[lab != 255], lab[lab != 255]
    assert len(np.unique(found)) == n_obj
    assert inertia(found) <= 1.02 * inertia(planted)


def test_attention_quarter_resolution_length(cuda):
    """L = 131 072 positions (the 1/4-resolution map of a 1024x2048 image), 2 heads, d = 12: the fused kernel never
    materialises L x L; 512 sampled query rows are compared with the dense fp64 softmax over all keys."""
    from isa_b200.attention import scaled_dot_product_attention
    torch.manual_seed(1)
    L, d = (H // 4) * (W // 4), 12
    q, k, v = [torch.randn(2, L, d, device=cuda) * 0.7 for _ in range(3)]
    v = v + 1.0          # zero-mean values would average to ~0 over 131 072 keys and make a relative error meaningless
    out, _ = scaled_dot_product_attention(q, k, v, d ** 0.5)
    idx = torch.randint(0, L, (512,), device=cuda)
    s = torch.einsum("bqd,bkd->bqk", q[:, idx].double(), k.double()) / d ** 0.5
    ref = torch.einsum("bqk,bkd->bqd", torch.softmax(s, dim=2), v.double())
    assert _rel(out[:, idx], ref) < 2e-4


def test_renet_quarter_resolution_map(cuda):
    """256 x 512 map (512-step row sweeps, 256-step column sweeps), n_units = 100: tensor-core scan against the nn.GRU
    oracle on the same weights (forward; the backward is covered at smaller sizes and by the wired-model test)."""
    from isa_b200.renet import ReNet
    from oracle.renet_ref import ReNetRef
    torch.manual_seed(3)
    ref = ReNetRef(16, 100)
    mod = ReNet(16, 100).to(cuda)
    mod.load_state_dict(ref.state_dict())
    x = torch.randn(1, 16, H // 4, W // 4)
    with torch.no_grad():
        y = mod(x.to(cuda))
        yr = ref(x)
    assert y.shape == (1, 200, H // 4, W // 4)
    assert _rel(y, yr) < 2e-4


def test_prediction_settings_and_upsampling(cuda):
    """The Cityscapes-shaped settings classes exist (the reference asserts CVPPP only, train.py:39) and the
    scatter / INTER_NEAREST kernels handle 1024x2048 -> 1024x2048 and -> 512x1024."""
    from isa_b200 import clustering
    from isa_b200.settings import CityscapesModelSettings
    ms = CityscapesModelSettings()
    assert (ms.IMAGE_HEIGHT, ms.IMAGE_WIDTH) == (H, W)
    rs = np.random.RandomState(0)
    sem = rs.uniform(size=(2, H, W)).astype(np.float32)
    emb = rs.standard_normal((4, H, W)).astype(np.float32)
    cls_map, Xt, fg_index, n_dev = clustering.fg_compact(torch.tensor(sem, device=cuda), torch.tensor(emb, device=cuda))
    fg_ref, X_ref = KM.gather_foreground(sem, emb)
    n = int(n_dev)
    assert n == len(X_ref) and np.array_equal(Xt[:, :n].t().cpu().numpy(), X_ref)
    labels = rs.randint(0, 64, size=H * W).astype(np.int32)
    small, up, cls_up = clustering.scatter_labels_upsample(torch.tensor(labels, device=cuda), fg_index, n_dev, cls_map, 512, 1024)
    ref_small = KM.scatter_labels(fg_ref, labels[:n])
    assert np.array_equal(small.cpu().numpy(), ref_small)
    assert np.array_equal(up.cpu().numpy(), KM.upsample_nearest(ref_small, 512, 1024))
