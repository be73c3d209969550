import os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
from isa_b200.model import Model
dev = torch.device("cuda:0")
torch.manual_seed(23)
model = Model('CVPPP', 'ReSeg', 2, 32, use_instance_segmentation=True, n_embedding=24, device=dev)
model.define_criterion(None, 0.5, 1.5, 2, False, 'Multi')
model.define_optimizer(1.0, 0.001, 0.5, 25, 'Adadelta')
img, sem, ins, labels, nobj = bench.train_batch(0, 16, "compact")
b = [torch.from_numpy(a).to(dev) for a in (img, sem, ins, nobj)]
for _ in range(4):
    model.train_step(b[0], b[1], b[2], b[3], 10.0)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    model.train_step(b[0], b[1], b[2], b[3], 10.0)
    torch.cuda.synchronize()
for e in prof.events():
    if e.name in ("aten::copy_", "aten::cat", "aten::add_", "aten::add", "aten::clone", "aten::contiguous", "aten::zeros", "aten::zero_", "aten::fill_", "aten::sum", "aten::mul") and e.device_time_total > 15:
        st = [s for s in (e.stack or []) if 'repo' in s][:3]
        print("%7.1f us %-16s %s | %s" % (e.device_time_total, e.name, str(e.input_shapes)[:70], " <- ".join(s.split('/')[-1] for s in st)))
