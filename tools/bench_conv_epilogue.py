"""Is cuDNN's fused conv + bias + ReLU faster than the plain convolution followed by the in-place epilogue kernel?
    python tools/bench_conv_epilogue.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isa_b200.pointwise import bias_act_  # noqa: E402

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


for (b, cin, cout, hw) in [(16, 3, 64, 256), (16, 64, 64, 256), (16, 64, 128, 128), (16, 128, 128, 128), (16, 128, 256, 64), (16, 256, 256, 64)]:
    x = torch.randn(b, cin, hw, hw, device=dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(cout, cin, 3, 3, device=dev) * 0.05).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(cout, device=dev)
    with torch.no_grad():
        t_plain = timed(lambda: F.conv2d(x, w, None, 1, 1))
        t_ours = timed(lambda: bias_act_(F.conv2d(x, w, None, 1, 1), bias, True))
        t_fused = timed(lambda: torch.cudnn_convolution_relu(x, w, bias, (1, 1), (1, 1), (1, 1), 1))
        y0 = bias_act_(F.conv2d(x, w, None, 1, 1), bias, True)
        y1 = torch.cudnn_convolution_relu(x, w, bias, (1, 1), (1, 1), (1, 1), 1)
        err = float((y0 - y1).abs().max() / y0.abs().max())
    print("conv %3d->%3d @%d: plain %.1f us, plain + epilogue kernel %.1f us, cuDNN fused %.1f us (rel diff %.1e, out channels_last %s)" % (
        cin, cout, hw, t_plain, t_ours, t_fused, err, y1.is_contiguous(memory_format=torch.channels_last)))
