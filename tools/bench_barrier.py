"""Cost of one grid-wide barrier (persistent cooperative kernels): python tools/bench_barrier.py"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isa_b200 import _lib  # noqa: E402

lib = _lib.load()
scratch = torch.zeros(4, dtype=torch.int32, device="cuda:0")
out = {}
for cps, thr in ((1, 256), (2, 256), (1, 512), (1, 1024)):
    for variant in (0, 1):
        us = ctypes.c_float(0)
        rc = lib.isa_selftest_grid_barrier(cps, thr, 2000, variant, scratch.data_ptr(), ctypes.byref(us))
        _lib.check(rc, "isa_selftest_grid_barrier")
        out["ctas/sm=%d threads=%d variant=%d" % (cps, thr, variant)] = round(us.value, 3)
for packed in (0, 1):
    for w in (4, 8, 16, 32, 64):
        r = ctypes.c_float(0)
        rc = lib.isa_selftest_fma_rate(packed, w, 1965.0, scratch.data_ptr(), ctypes.byref(r))
        _lib.check(rc, "isa_selftest_fma_rate")
        out["fma/clk/SM %s warps=%d (at 1965 MHz)" % ("FFMA2" if packed else "FFMA", w)] = round(r.value, 1)
for cols in (16, 32):
    for w in (4, 8, 16):
        r = ctypes.c_float(0)
        rc = lib.isa_selftest_tmem_ld_rate(w, cols, 1965.0, scratch.data_ptr(), ctypes.byref(r))
        _lib.check(rc, "isa_selftest_tmem_ld_rate")
        out["tmem ld B/clk/SM x%d warps=%d (at 1965 MHz)" % (cols, w)] = round(r.value, 1)
print(json.dumps(out, indent=1))
