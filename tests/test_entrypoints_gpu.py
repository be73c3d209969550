"""GPU: the reference-named entry points run end to end (train -> checkpoint -> pred_list -> evaluate file contract)."""
import os
import subprocess
import sys

import numpy as np
import pytest
from PIL import Image

from isa_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, cwd):
    r = subprocess.run([sys.executable] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_train_pred_evaluate_roundtrip(tmp_path):
    out = str(tmp_path)
    _run([os.path.join(ROOT, "code", "train.py"), "--nepochs", "1", "--batchsize", "2", "--synthetic", "2", "--output", os.path.join(out, "models")], ROOT)
    ckpts = [f for f in os.listdir(os.path.join(out, "models", "CVPPP")) if f.endswith(".pth")]
    assert ckpts and os.path.exists(os.path.join(out, "models", "CVPPP", "training.log"))
    gt_dir = os.path.join(out, "gt")
    os.makedirs(gt_dir)
    names = []
    for j in range(2):
        nm = "plant%03d_rgb" % j
        Image.fromarray(synth.leaf_image(j, 120, 110)).save(os.path.join(out, nm + ".png"))
        lab = synth.label_map(np.random.RandomState(j), 120, 110, 4).astype(np.int64) + 1
        lab[lab == 256] = 0
        Image.fromarray(lab.astype(np.uint8)).save(os.path.join(gt_dir, "plant%03d_label.png" % j))
        Image.fromarray((lab > 0).astype(np.uint8)).save(os.path.join(gt_dir, "plant%03d_fg.png" % j))
        names.append(os.path.join(out, nm + ".png"))
    with open(os.path.join(out, "val.lst"), "w") as f:
        f.write("\n".join(names) + "\n")
    with open(os.path.join(out, "counts.txt"), "w") as f:
        f.write("plant000,4\nplant001,4\n")
    _run([os.path.join(ROOT, "code", "pred_list.py"), "--lst", os.path.join(out, "val.lst"), "--model",
          os.path.join(out, "models", "CVPPP", ckpts[0]), "--output", os.path.join(out, "pred")], ROOT)
    for nm in ("plant000_rgb", "plant001_rgb"):
        d = os.path.join(out, "pred", nm)
        for suffix in (".png", "-fg_mask.png", "-ins_mask.png", "-ins_mask_color.png", "-n_objects.npy"):
            assert os.path.exists(os.path.join(d, nm + suffix)), suffix
        assert np.array(Image.open(os.path.join(d, nm + "-ins_mask.png"))).shape == (120, 110)
    txt = _run([os.path.join(ROOT, "code", "evaluate.py"), "--pred_dir", os.path.join(out, "pred"), "--dataset", "CVPPP",
                "--names", os.path.join(out, "val.lst"), "--counts", os.path.join(out, "counts.txt"), "--gt_dir", gt_dir], ROOT)
    assert "MEAN SBD" in txt and "MEAN |DIC|" in txt
    _run([os.path.join(ROOT, "code", "pred.py"), "--image", names[0], "--output", os.path.join(out, "single")], ROOT)
    assert os.path.exists(os.path.join(out, "single", "plant000_rgb-ins_mask.png"))
