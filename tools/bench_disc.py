"""Micro-benchmark of the discriminative-loss kernels through the raw C-ABI (preallocated
buffers, CUDA events on the launching stream, L2 flushed between timed launches).
python tools/bench_disc.py [bs C H W K]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isa_b200 import _lib, synth  # noqa: E402


def time_kernel(fn, iters=30, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
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
    lib = _lib.load()
    dev = torch.device("cuda:0")
    d = synth.batch(0, bs, C, H, W, K)
    x = torch.tensor(d["emb"], device=dev)
    n = torch.tensor(d["n_objects"], device=dev, dtype=torch.int32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss = torch.empty(1, device=dev)
    terms = torch.empty(4, device=dev)
    means = torch.empty(bs, K, C, device=dev)
    grad = torch.empty_like(x)
    gl = torch.ones(1, device=dev)
    wsb = lib.isa_disc_loss_workspace_bytes(bs, C, K, H, W)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    res = {"shape": [bs, C, H, W, K]}
    P = H * W
    for kind, code, tgt, tb in (("label_u8", 0, torch.tensor(d["labels"], device=dev), 1),
                                ("dense_f32", 1, torch.tensor(synth.onehot(d["labels"], K), device=dev), 4 * K),
                                ("dense_i64", 2, torch.tensor(synth.onehot(d["labels"], K, np.int64), device=dev), 8 * K)):
        def fwd():
            rc = lib.isa_disc_loss_fwd(x.data_ptr(), tgt.data_ptr(), code, n.data_ptr(), bs, C, H, W, K, 0.5, 1.5, 2, 1,
                                       1.0, 0.0, 0.0, 0.005, None, loss.data_ptr(), terms.data_ptr(), means.data_ptr(),
                                       ws.data_ptr(), wsb, st)
            assert rc == 0, lib.isa_last_error()

        def bwd():
            rc = lib.isa_disc_loss_bwd(x.data_ptr(), tgt.data_ptr(), code, n.data_ptr(), bs, C, H, W, K, 0.5, 1.5, 2, 1,
                                       1.0, 0.0, 0.0, 0.005, None, means.data_ptr(), gl.data_ptr(), None, grad.data_ptr(),
                                       ws.data_ptr(), wsb, st)
            assert rc == 0, lib.isa_last_error()

        f_med, f_min = time_kernel(fwd, flush=flush)
        b_med, b_min = time_kernel(bwd, flush=flush)
        fbytes = bs * P * (4 * C + tb) + bs * K * C * 4
        bbytes = bs * P * (8 * C + tb)
        res[kind] = {"fwd_us": round(f_med, 1), "fwd_min_us": round(f_min, 1), "fwd_GBs": round(fbytes / f_med / 1e3),
                     "bwd_us": round(b_med, 1), "bwd_min_us": round(b_min, 1), "bwd_GBs": round(bbytes / b_med / 1e3)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
