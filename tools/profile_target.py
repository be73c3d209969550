"""Short deterministic workload for ncu: python tools/profile_target.py [disc|kmeans|all]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isa_b200 import clustering, synth  # noqa: E402
from isa_b200.losses import DiscriminativeLoss  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
if what in ("disc", "all"):
    d = synth.batch(0, 16, 24, 256, 256, 32)
    x = torch.tensor(d["emb"], device=dev, requires_grad=True)
    lab = torch.tensor(d["labels"], device=dev)
    n = torch.tensor(d["n_objects"], device=dev)
    crit = DiscriminativeLoss(0.5, 1.5, 2)
    for _ in range(3):
        loss, _ = crit(x, lab, n, 32)
        loss.backward()
    torch.cuda.synchronize()
    print("disc loss", float(loss))
if what in ("kmeans", "all"):
    rs = np.random.RandomState(0)
    k, C, nn = 16, 24, 40000
    cent = rs.standard_normal((k, C))
    cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    X = (0.3 * rs.standard_normal((nn, C)) + 0.7 * cent[rs.randint(0, k, nn)]).astype(np.float32)
    Xt = torch.tensor(X, device=dev).t().contiguous()
    n_dev = torch.tensor([nn], device=dev, dtype=torch.int32)
    for _ in range(2):
        res = clustering.kmeans_fit(Xt, n_dev, k, seed=0, n_init=35)
    torch.cuda.synchronize()
    print("kmeans iters", int(res.n_iter.sum()))
if what in ("renet",):
    from isa_b200.renet import ReNet
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    mod = ReNet(256, 100).to(dev)
    x = torch.randn(B, 256, 64, 64, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    for _ in range(2):
        y = mod(x)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
    print("renet", float(y.sum()))
if what in ("attn",):
    from isa_b200.attention import scaled_dot_product_attention
    torch.manual_seed(0)
    q, k, v = [torch.randn(32, 4096, 12, device=dev, requires_grad=True) for _ in range(3)]
    for _ in range(2):
        o, _ = scaled_dot_product_attention(q, k, v, 12 ** 0.5)
    torch.cuda.synchronize()
    print("attn", float(o.sum()))
if what in ("train",):
    from isa_b200.model import Model
    from isa_b200.settings import CVPPPTrainingSettings
    sys.path.insert(0, ROOT)
    import bench
    ts = CVPPPTrainingSettings()
    torch.manual_seed(23)
    model = Model('CVPPP', 'ReSeg', 2, 32, use_instance_segmentation=True, n_embedding=24, device=dev)
    model.define_criterion(None, 0.5, 1.5, 2, False, 'Multi')
    model.define_optimizer(1.0, 0.001, 0.5, 25, 'Adadelta')
    img, sem, ins, labels, nobj = bench.train_batch(0, 16)
    b = [torch.from_numpy(a).to(dev) for a in (img, sem, ins, nobj)]
    for _ in range(3):
        m = model.train_step(b[0], b[1], b[2], b[3], 10.0)
    torch.cuda.synchronize()
    print("train", float(m['Cost']))
