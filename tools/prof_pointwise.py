"""1x1 conv (conv_tc_kernel) time vs pixel count: separates the fixed per-launch / per-tile latency from throughput.
    python tools/prof_pointwise.py"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'ofa-for-super-resolution_b200'))
from ofa_b200 import functional as OF, backend as B
dev = torch.device('cuda:0')
for cin, cout in ((64, 384), (384, 64), (64, 64)):
    w = torch.randn(cout, cin, 1, 1, device=dev) * 0.05
    for n in (16, 64, 128, 256, 1024):
        x = torch.randn(n, cin, 24, 24, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        cache = OF.PackedWeightCache()
        f = lambda: OF.conv_bn_act_infer(x, w, cin, cout, 1, None, B.ACT_NONE, cache=cache, out_dtype=torch.bfloat16)
        for _ in range(3):
            f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        P = n * 576
        print(f'{cin:3d}->{cout:3d}  P={P:7d} ({P // 256:5d} tiles)  {us:7.1f} us   {P * (cin + cout) * 2 / us / 1e3:7.0f} GB/s')
