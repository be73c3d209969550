"""A few training steps for ncu launch lists / captures (cuDNN's algorithm search runs in the first ones; the LAST step
is steady state): python tools/ncu_step.py [batch] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from isa_b200.model import Model  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
torch.manual_seed(23)
model = Model('CVPPP', 'ReSeg', 2, 32, use_instance_segmentation=True, n_embedding=24, device=dev)
model.define_criterion(None, 0.5, 1.5, 2, False, 'Multi')
model.define_optimizer(1.0, 0.001, 0.5, 25, 'Adadelta')
img, sem, ins, labels, nobj = bench.train_batch(0, bs, "compact")
b = [torch.from_numpy(a).to(dev) for a in (img, sem, ins, nobj)]
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    m = model.train_step(b[0], b[1], b[2], b[3], 10.0)
torch.cuda.synchronize()
print("cost", float(m['Cost']))
