"""Micro-benchmark of the discriminative-loss kernels (CUDA events, L2 flushed between
iterations).  python tools/bench_disc.py [bs C H W K]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isa_b200 import synth  # noqa: E402
from isa_b200.losses import DiscriminativeLoss  # noqa: E402


def timeit(fn, iters=20, warm=5, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    bs, C, H, W, K = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else (16, 24, 256, 256, 32)
    dev = torch.device("cuda:0")
    d = synth.batch(0, bs, C, H, W, K)
    x = torch.tensor(d["emb"], device=dev, requires_grad=True)
    n = torch.tensor(d["n_objects"], device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    crit = DiscriminativeLoss(0.5, 1.5, 2)
    res = {"shape": [bs, C, H, W, K]}
    P = H * W
    for kind, tgt, tb in (("label_u8", torch.tensor(d["labels"], device=dev), 1),
                          ("dense_f32", torch.tensor(synth.onehot(d["labels"], K), device=dev), 4 * K),
                          ("dense_i64", torch.tensor(synth.onehot(d["labels"], K, np.int64), device=dev), 8 * K)):
        state = {}

        def fwd():
            state["loss"], _ = crit(x, tgt, n, K)

        def bwd():
            state["loss"].backward(retain_graph=True)

        f_med, f_min = timeit(fwd, flush=flush)
        fwd()
        b_med, b_min = timeit(bwd, flush=flush)
        fbytes = bs * P * (4 * C + tb) + bs * K * C * 4
        bbytes = bs * P * (8 * C + tb)
        res[kind] = {"fwd_us": f_med, "fwd_min_us": f_min, "fwd_GBs": fbytes / f_med / 1e3,
                     "bwd_us": b_med, "bwd_min_us": b_min, "bwd_GBs": bbytes / b_med / 1e3}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
