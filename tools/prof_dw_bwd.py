"""Times the depthwise backward kernels (filter gradient, data gradient) at the C3 training shape
(batch 64 of 24x24 LR patches, bf16 NHWC).  Usage: python tools/prof_dw_bwd.py [C] [N] [H] [W]"""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'ofa-for-super-resolution_b200'))
from ctypes import byref
from ofa_b200 import backend as B

a = [int(v) for v in sys.argv[1:]]
C = a[0] if len(a) > 0 else 384
N = a[1] if len(a) > 1 else 64
H = a[2] if len(a) > 2 else 24
W = a[3] if len(a) > 3 else 24
dev = torch.device('cuda:0')
x = torch.randn(N, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
dy = torch.randn(N, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
L = B.lib()
st = B.stream_ptr(dev)
for ks in (3, 5, 7):
    dwa = torch.empty(C, ks * ks, device=dev)
    tx, tdy = B.t4(x), B.t4(dy)
    f = lambda: B.check(L.ofa_dw_bwd_filter(byref(tx), byref(tdy), ks, dwa.data_ptr(), st))
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    P = N * H * W
    print(f'dw_bwd_filter ks={ks} C={C} P={P}: {us:.1f} us  ({2 * P * C * ks * ks / us / 1e6:.2f} TFLOP/s, '
          f'{2 * P * C * 2 / us / 1e3:.0f} GB/s algorithmic)')
