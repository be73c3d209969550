"""pred.py inference on the benchmark images for ncu captures and the Lloyd kernel's own phase profile:
    python tools/ncu_infer.py [n_images] [dump_dir]
Prints, per image, fg points, restart-iterations, grid iterations and CTA 0's phase times (E-step, barrier, update,
barrier; us).  With dump_dir the softmax map and the embedding of every image are saved for CPU-side parity analysis."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from isa_b200 import settings, synth  # noqa: E402
from isa_b200.model import Model  # noqa: E402
from isa_b200.prediction import Prediction  # noqa: E402

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dump = sys.argv[2] if len(sys.argv) > 2 else None
dev = torch.device("cuda:0")
torch.manual_seed(23)
model = Model('CVPPP', 'ReSeg', 2, 32, use_instance_segmentation=True, n_embedding=24, device=dev)
ms_ = settings.CVPPPModelSettings()
pred = Prediction(ms_.IMAGE_HEIGHT, ms_.IMAGE_WIDTH, ms_.MEAN, ms_.STD, False, model, 1, seed=0)
for j in range(n_img):
    raw = synth.leaf_image(j, bench.RAW_H, bench.RAW_W)
    t = pred.image_to_tensor(raw)[0].unsqueeze(0).to(dev)
    sem, emb = model.predict_device(t)
    out = pred.cluster_device(sem[0], emb[0], bench.N_OBJ, bench.RAW_H, bench.RAW_W)
    torch.cuda.synchronize()
    res = out[4]
    info = res.info.cpu().numpy() if hasattr(res.info, "cpu") else np.asarray(res.info)
    print("image %d: n=%d restart-iterations=%d grid-iterations=%d phases(us) E=%d bar1=%d upd=%d bar2=%d loop=%d tail=%d | seeding(us) search=%d pass1=%d pass2=%d barriers=%d segments=%d rounds=%d" % (
        j, int(info[2]), int(res.n_iter.sum()), int(info[3]), info[4], info[5], info[6], info[7], info[8], info[9], info[10], info[11], info[12], info[13], info[14], info[15]))
    if dump:
        os.makedirs(dump, exist_ok=True)
        from oracle import kmeans as KM   # diagnostic tool (not product): the oracle's own foreground rule
        fg, X = KM.gather_foreground(sem[0].cpu().numpy(), emb[0].cpu().numpy())
        np.savez_compressed(os.path.join(dump, "X_img%d.npz" % j), fg=fg, X=X)
