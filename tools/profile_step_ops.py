"""Operator-level (aten) device-time breakdown of one training step with input shapes and the autograd node that
launched them: python tools/profile_step_ops.py [batch]"""
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
for _ in range(4):
    model.train_step(b[0], b[1], b[2], b[3], 10.0)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    model.train_step(b[0], b[1], b[2], b[3], 10.0)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = getattr(e, "self_device_time_total", 0)
    if t > 0:
        rows.append((t, e.count, e.key, str(e.input_shapes)[:150]))
rows.sort(reverse=True)
print("self device time: %.2f ms" % (sum(r[0] for r in rows) / 1e3))
for t, c, k, s in rows[:70]:
    print("%8.1f us %4d x %-44s %s" % (t, c, k[:44], s))
