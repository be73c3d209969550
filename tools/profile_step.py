"""Steady-state device-time breakdown of one training step (torch.profiler): python tools/profile_step.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from isa_b200.model import Model  # noqa: E402
from isa_b200.settings import CVPPPTrainingSettings  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
ts = CVPPPTrainingSettings()
torch.manual_seed(23)
model = Model('CVPPP', 'ReSeg', 2, 32, use_instance_segmentation=True, n_embedding=24, device=dev)
model.define_criterion(None, 0.5, 1.5, 2, False, 'Multi')
model.define_optimizer(1.0, 0.001, 0.5, 25, 'Adadelta')
img, sem, ins, labels, nobj = bench.train_batch(0, bs, "compact")
b = [torch.from_numpy(a).to(dev) for a in (img, sem, ins, nobj)]
for _ in range(4):
    model.train_step(b[0], b[1], b[2], b[3], 10.0)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        model.train_step(b[0], b[1], b[2], b[3], 10.0)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    if t > 0 and e.device_type.name == "CUDA" if hasattr(e, "device_type") else t > 0:
        rows.append((t / 3.0, e.count // 3, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("device time per step: %.2f ms over %d kernel kinds" % (tot / 1e3, len(rows)))
for t, c, k in rows[:45]:
    print("%8.1f us  %4d x  %s" % (t, c, k[:120]))
