"""Micro-benchmark of a ReNet layer: python tools/bench_renet.py [B C H W n]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isa_b200.renet import ReNet  # noqa: E402


def main():
    B, C, H, W, n = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else (16, 256, 64, 64, 100)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    mod = ReNet(C, n).to(dev)
    x = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    res = {"shape": [B, C, H, W, n]}

    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.median(ts))

    with torch.no_grad():
        res["fwd_infer_ms"] = timeit(lambda: mod(x))
    st = {}

    def f():
        st["y"] = mod(x)

    res["fwd_train_ms"] = timeit(f)
    f()
    g = torch.randn_like(st["y"])
    res["bwd_ms"] = timeit(lambda: st["y"].backward(g, retain_graph=True))
    # cuDNN nn.GRU for the same two sweeps
    rows = torch.randn(B * H, W, C, device=dev)
    gru1 = torch.nn.GRU(C, n, batch_first=True, bidirectional=True).to(dev)
    gru2 = torch.nn.GRU(2 * n, n, batch_first=True, bidirectional=True).to(dev)
    cols = torch.randn(B * W, H, 2 * n, device=dev)
    with torch.no_grad():
        res["cudnn_fwd_infer_ms"] = timeit(lambda: (gru1(rows), gru2(cols)))
    # per-kernel device times of the scan entry points (CUDA events on the launching stream)
    from isa_b200 import _lib
    _lib.TIMER.enabled = True
    _lib.TIMER.reset()
    for _ in range(5):
        f()
        st["y"].backward(g)
    for name, (cnt, ms) in _lib.TIMER.summary().items():
        res[name + "_us_per_launch"] = 1e3 * ms / cnt
    _lib.TIMER.enabled = False
    print(json.dumps(res))


if __name__ == "__main__":
    main()
