"""Micro-benchmark of the attention kernels: python tools/bench_attn.py [BH L d]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isa_b200 import _lib  # noqa: E402


def main():
    BH, L, d = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else (32, 4096, 12)
    lib = _lib.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    q, k, v = [torch.randn(BH, L, d, device=dev) for _ in range(3)]
    out = torch.empty_like(q)
    lse = torch.empty(BH, L, device=dev)
    wsb = lib.isa_attention_workspace_bytes(BH, L, L)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    T = float(np.sqrt(d))
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    go = torch.randn_like(out)

    def fwd():
        rc = lib.isa_attention_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), BH, L, L, d, d, T, None, None, 1,
                                   out.data_ptr(), lse.data_ptr(), ws.data_ptr(), wsb, st)
        assert rc == 0

    def bwd():
        rc = lib.isa_attention_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), go.data_ptr(), lse.data_ptr(),
                                   BH, L, L, d, d, T, None, None, 1, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), wsb, st)
        assert rc == 0

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

    f = timeit(fwd)
    b = timeit(bwd)
    flops = 4.0 * BH * L * L * d
    res = {"BH": BH, "L": L, "d": d, "fwd_ms": f, "bwd_ms": b, "fwd_TFLOPs_algo": flops / f / 1e9,
           "fwd_Gexp_per_s": BH * L * L / f / 1e6}
    # library baselines on the same problem
    qq, kk, vv = [t.view(1, BH, L, d) for t in (q, k, v)]
    res["torch_sdpa_fp32_ms"] = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv))
    qb, kb, vb = [t.bfloat16() for t in (qq, kk, vv)]
    res["torch_sdpa_bf16_ms"] = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qb, kb, vb))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
