"""Times the three ReNet projection GEMMs (csrc/proj_gemm.cu) in isolation: python tools/bench_proj.py [tokens cin n]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isa_b200 import _lib  # noqa: E402

tokens, cin, n = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else (65536, 256, 100)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
lib = _lib.load()
dev = torch.device("cuda:0")
x = torch.randn(tokens, cin, device=dev)
w = torch.randn(6 * n, cin, device=dev)
gx = torch.empty(tokens, 6 * n, device=dev)
dg = torch.randn(tokens, 2, 3 * n, device=dev)
dghn = torch.randn(tokens, 2, n, device=dev)
out = torch.randn(tokens, 2, n, device=dev)
dx = torch.empty(tokens, cin, device=dev)
dw_ih, dw_hh = torch.empty(2, 3 * n, cin, device=dev), torch.empty(2, 3 * n, n, device=dev)
db_ih, db_hh = torch.empty(2, 3 * n, device=dev), torch.empty(2, 3 * n, device=dev)
wsb = lib.isa_renet_proj_wgrad_workspace_bytes(tokens, cin, n)
ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
wpb = lib.isa_renet_proj_workspace_bytes(cin, 6 * n)
wpk = torch.empty(wpb, device=dev, dtype=torch.uint8)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
P = _lib.ptr


def fwd():
    _lib.check(lib.isa_renet_proj_fwd(P(x), P(w), tokens, cin, 6 * n, P(gx), P(wpk), wpb, st), "fwd")


def dxf():
    _lib.check(lib.isa_renet_proj_dx(P(dg), P(w), tokens, 6 * n, cin, P(dx), P(wpk), wpb, st), "dx")


def wg():
    _lib.check(lib.isa_renet_proj_wgrad(P(dg), P(dghn), P(x), P(out), tokens, cin, n, 1, 1, 64, P(dw_ih), P(dw_hh), P(db_ih), P(db_hh), P(ws), wsb, st), "wgrad")


res = {"shape": [tokens, cin, n], "wgrad_workspace_MB": wsb / 1e6}
for name, fn, flops, byts in (("fwd", fwd, 2.0 * tokens * 6 * n * cin, 4.0 * tokens * (cin + 6 * n)),
                              ("dx", dxf, 2.0 * tokens * 6 * n * cin, 4.0 * tokens * (cin + 6 * n)),
                              ("wgrad", wg, 2.0 * tokens * (6 * n * cin + 6 * n * n), 4.0 * tokens * (cin + 10 * n))):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    t = float(np.median(ts))
    res[name] = {"us": round(t, 1), "useful_TFLOPs": round(flops / t / 1e6, 1), "issued_bf16_TFLOPs": round(3 * flops / t / 1e6, 1),
                 "algorithmic_GBs": round(byts / t / 1e3)}
print(json.dumps(res))
