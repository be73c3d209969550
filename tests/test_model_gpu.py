"""GPU: the wired hot path (EmbeddingPath / ReSeg / Model / Prediction) against the CPU composition of
reference ops (oracle/model_ref.py) on identical weights and inputs.
  * embeddings: <= 1e-3 relative (north_star); asserted at 5e-4
  * loss: the reference's own broadcast graph restated in torch (checked against the golden-pinned numpy oracle)
  * instance masks: bit-exact against the oracle pipeline; SBD / |DiC| identical through the evaluate.py functions."""
import numpy as np
import pytest
import torch

from isa_b200 import metrics, synth
from oracle import disc_loss as O
from oracle import evaluate_ref as E
from oracle import kmeans as KM
from oracle.model_ref import EmbeddingPathRef, discriminative_loss_torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_reference_loss_graph_matches_numpy_oracle():
    d = synth.batch(1, 2, 8, 12, 12, 4, n_min=1, n_max=4)
    x = torch.tensor(d["emb"], requires_grad=True)
    loss, means = discriminative_loss_torch(x, torch.tensor(synth.onehot(d["labels"], 4)), d["n_objects"], 4)
    loss.backward()
    o = O.discriminative_loss(d["emb"], d["labels"], d["n_objects"], 4, 0.5, 1.5, 2, want_grad=True)
    assert abs(float(loss) - float(o["loss"])) < 1e-5
    assert np.abs(x.grad.numpy() - o["grad"]).max() < 1e-5 * np.abs(o["grad"]).max() + 1e-8


def test_embedding_path_matches_reference_composition(cuda):
    from isa_b200.archs import EmbeddingPath
    torch.manual_seed(1)
    ref = EmbeddingPathRef(32, 20).double()
    mod = EmbeddingPath(32, 20).to(cuda).eval()
    mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    x = torch.randn(2, 32, 12, 16, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    xg = x.float().to(cuda).requires_grad_(True)
    yg = mod(xg)
    assert _rel(yg, yr) < 5e-4
    yg.backward(gy.float().to(cuda))
    assert _rel(xg.grad, xr.grad) < 5e-4
    gmax = max(float(q.grad.abs().max()) for q in ref.parameters())
    for (name, p), (_, q) in zip(mod.named_parameters(), ref.named_parameters()):
        # w_ks.bias has an exactly-zero gradient (softmax shift invariance): floor the scale
        scale = max(float(q.grad.abs().max()), 1e-2 * gmax)
        assert float((p.grad.double().cpu() - q.grad).abs().max()) < 1e-3 * scale, name


def _make_model(cuda, K=32, C=24):
    from isa_b200.model import Model
    torch.manual_seed(23)
    m = Model('CVPPP', 'ReSeg', 2, K, use_instance_segmentation=True, n_embedding=C, device=cuda)
    return m


def test_train_step_runs_and_learns(cuda):
    m = _make_model(cuda)
    m.define_criterion(None, 0.5, 1.5, 2, False, 'Multi')
    m.define_optimizer(1.0, 0.001, 0.5, 25, 'Adadelta')
    d = synth.batch(5, 2, 3, 64, 64, 32, n_min=2, n_max=5)
    images = torch.tensor(d["emb"])
    sem = torch.tensor(np.stack([(d["labels"] == 255), (d["labels"] != 255)], 1).astype(np.int64))
    ins = torch.tensor(synth.onehot(d["labels"], 32, np.int64))
    costs = []
    for _ in range(8):
        out = m.train_step(images, sem, ins, torch.tensor(d["n_objects"]), 10.0)
        costs.append(float(out['Cost']))
    assert all(np.isfinite(costs)), costs
    assert costs[-1] < costs[0], costs
    # label-map targets give the same step
    m2 = _make_model(cuda)
    m2.define_criterion(None, 0.5, 1.5, 2, False, 'Multi')
    m2.define_optimizer(1.0, 0.001, 0.5, 25, 'Adadelta')
    a = m2.train_step(images, sem, ins, torch.tensor(d["n_objects"]), 10.0)
    m3 = _make_model(cuda)
    m3.define_criterion(None, 0.5, 1.5, 2, False, 'Multi')
    m3.define_optimizer(1.0, 0.001, 0.5, 25, 'Adadelta')
    b = m3.train_step(images, sem, torch.tensor(d["labels"]), torch.tensor(d["n_objects"]), 10.0)
    assert abs(float(a['Cost']) - float(b['Cost'])) < 1e-5 * abs(float(a['Cost']))


def test_loss_of_model_output_matches_reference_graph(cuda):
    from isa_b200.losses import DiscriminativeLoss
    m = _make_model(cuda)
    d = synth.batch(6, 2, 3, 32, 32, 32, n_min=2, n_max=6)
    with torch.no_grad():
        _, emb = m.model(False, torch.tensor(d["emb"], device=cuda))
    emb = emb.detach().requires_grad_(True)
    tgt = torch.tensor(synth.onehot(d["labels"], 32))
    loss, means = DiscriminativeLoss(0.5, 1.5, 2)(emb, tgt.to(cuda), torch.tensor(d["n_objects"]), 32)
    loss.backward()
    e2 = emb.detach().cpu().requires_grad_(True)
    lref, mref = discriminative_loss_torch(e2, tgt, d["n_objects"], 32)
    lref.backward()
    assert abs(float(loss) - float(lref)) < 1e-4 * abs(float(lref))
    assert _rel(means, mref) < 1e-4
    assert _rel(emb.grad, e2.grad) < 1e-4


def test_prediction_masks_bit_exact_and_sbd_identical(cuda, tmp_path):
    from isa_b200.prediction import Prediction
    from isa_b200.settings import CVPPPModelSettings
    ms = CVPPPModelSettings()
    m = _make_model(cuda)
    pred = Prediction(64, 64, ms.MEAN, ms.STD, False, m, 1, seed=0, n_init=6)
    m.n_objects_prediction = 5
    raw = synth.leaf_image(2, 133, 125)
    sem_seg, ins_seg, n = pred.predict_array(raw)
    assert sem_seg.shape == (133, 125) and ins_seg.shape == (133, 125) and sem_seg.dtype == np.uint8
    # oracle pipeline on the SAME network outputs (taken from the device): sklearn-rule k-means oracle,
    # numpy scatter, cv2 nearest up-sampling
    image, h, w = pred.image_to_tensor(raw)
    sem_p, emb = m.predict_device(image.unsqueeze(0))
    semn, embn = sem_p[0].cpu().numpy(), emb[0].cpu().numpy()
    fg, X = KM.gather_foreground(semn, embn)
    if len(X) >= 5:
        o = KM.kmeans_oracle(X, 5, seed=0, n_init=6)
        mask = KM.scatter_labels(fg, o["labels"])
        want_ins = KM.upsample_nearest(mask, h, w)
        want_sem = KM.upsample_nearest(fg, h, w)
        assert np.array_equal(ins_seg, want_ins)
        assert np.array_equal(sem_seg, want_sem)
        gt = synth.label_map(np.random.RandomState(0), 133, 125, 5).astype(np.int64) + 1
        gt[gt == 256] = 0
        if (ins_seg > 0).any():
            assert metrics.calc_sbd(gt, ins_seg) == E.calc_sbd(gt, want_ins)
            assert metrics.calc_dic(5, len(np.unique(ins_seg)) - 1) == E.calc_dic(5, len(np.unique(want_ins)) - 1)
    # file API + reference return contract
    from PIL import Image
    p = str(tmp_path / "img.png")
    Image.fromarray(raw).save(p)
    r, s, i, n2 = pred.predict(p)
    assert np.array_equal(r, raw) and np.array_equal(i, ins_seg) and n2 == 5
    sem_t, ins_t, n_t = m.predict(image.unsqueeze(0))
    assert sem_t.shape == (1, 2, 64, 64) and ins_t.shape == (1, 24, 64, 64) and n_t.tolist() == [[5]]
    a, b, c = pred.cluster(sem_t[0], ins_t[0], n_t[0])
    assert a.dtype == np.uint8 and b.shape == (64, 64)


def test_predict_many_equals_predict_array_and_prefetcher_is_lossless(cuda):
    """The pipelined public paths (Prediction.predict_many, data.CudaPrefetcher) change scheduling, not results."""
    from isa_b200 import synth
    from isa_b200.data import CudaPrefetcher
    from isa_b200.model import Model
    from isa_b200.prediction import Prediction
    from isa_b200.settings import CVPPPModelSettings
    torch.manual_seed(3)
    ms = CVPPPModelSettings()
    model = Model('CVPPP', 'ReSeg', 2, 32, use_instance_segmentation=True, n_embedding=24, device=cuda)
    pred = Prediction(64, 64, ms.MEAN, ms.STD, False, model, 1, seed=0, n_init=3)
    model.n_objects_prediction = 4
    raws = [synth.leaf_image(s, 90 + 7 * s, 80 + 5 * s) for s in range(4)]
    many = list(pred.predict_many(iter(raws)))
    assert len(many) == len(raws)
    for raw, (sem, ins, n) in zip(raws, many):
        sem1, ins1, n1 = pred.predict_array(raw)
        assert n == n1 and np.array_equal(sem, sem1) and np.array_equal(ins, ins1)
    batches = [(torch.randn(2, 3, 8, 8), torch.randint(0, 5, (2, 4), dtype=torch.int64), torch.tensor([1, 2])) for _ in range(5)]
    got = list(CudaPrefetcher(batches, cuda))
    assert len(got) == 5
    for b, g in zip(batches, got):
        for t, u in zip(b, g):
            assert u.is_cuda and torch.equal(t, u.cpu())


def test_loss_forward_is_bit_reproducible(cuda):
    """The label-map forward of the discriminative loss (one grid barrier, per-warp accumulators folded in a fixed order, no
    float atomics) and the fused CE + Dice return bit-identical values run after run."""
    from isa_b200.losses import DiscriminativeLoss
    from isa_b200.seg_losses import SegLosses
    d = synth.batch(3, 16, 24, 256, 256, 32)
    x = torch.tensor(d["emb"], device=cuda)
    lab = torch.tensor(d["labels"], device=cuda)
    nobj = torch.tensor(d["n_objects"], device=cuda)
    crit = DiscriminativeLoss(0.5, 1.5, 2)
    vals, means = [], []
    for _ in range(4):
        l, m = crit(x, lab, nobj, 32)
        vals.append(float(l)); means.append(m.clone())
    assert len(set(vals)) == 1, vals
    assert all(torch.equal(means[0], m) for m in means[1:])
    z = torch.randn(16, 2, 256, 256, device=cuda)
    cm = (lab != 255).to(torch.uint8)
    outs = [tuple(float(v) for v in SegLosses()(z, cm, time=1)) for _ in range(3)]
    assert len(set(outs)) == 1, outs


def test_cuda_graph_replay_equals_eager_steps(cuda):
    """Model.enable_cuda_graph(): the captured step replays the same arithmetic as the eager step.  The FIRST step's cost is
    compared bit for bit (the forward -- scans, attention, heads, the losses with their fixed-order reductions -- has no
    float atomics); later steps agree to fp32 noise (the attention / scan backward kernels accumulate a few gradients with
    float atomics, so parameters differ in the last bits after an update)."""
    batches = []
    for seed in (5, 6):
        d = synth.batch(seed, 2, 3, 64, 64, 32, n_min=2, n_max=5)
        batches.append([torch.tensor(d["emb"], device=cuda),
                        torch.tensor(np.stack([(d["labels"] == 255), (d["labels"] != 255)], 1).astype(np.int64), device=cuda),
                        torch.tensor(synth.onehot(d["labels"], 32, np.int64), device=cuda),
                        torch.tensor(d["n_objects"], device=cuda)])
    runs = []
    for graph in (False, True):
        m = _make_model(cuda)
        m.define_criterion(None, 0.5, 1.5, 2, False, 'Multi')
        m.define_optimizer(1.0, 0.001, 0.5, 25, 'Adadelta')
        if graph:
            m.enable_cuda_graph(warmup_steps=2)
        costs = []
        for i in range(6):
            b = batches[i % 2]
            costs.append(float(m.train_step(b[0], b[1], b[2], b[3], 10.0)['Cost']))
        if graph:
            assert len(m._graphs) == 1
        runs.append((costs, torch.cat([p.detach().reshape(-1) for p in m.model.parameters()]).clone()))
    (c0, p0), (c1, p1) = runs
    assert c1[0] == c0[0], (c1[0], c0[0])
    np.testing.assert_allclose(c1, c0, rtol=2e-4)
    assert _rel(p1, p0) < 1e-3
