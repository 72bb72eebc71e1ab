"""Runs the individual hot kernels at the bench workload's shapes (for ncu captures).
    python tools/prof_kernels.py [expand|project|dw7|ps|out|all]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200'))
import torch
from ofa_b200 import functional as OF, backend as B

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
dev = torch.device('cuda:0')
H, W = 540, 960
torch.manual_seed(0)


class BN:
    def __init__(self, c):
        self.weight = torch.rand(c, device=dev) + 0.5
        self.bias = torch.randn(c, device=dev) * 0.1
        self.running_mean = torch.randn(c, device=dev) * 0.1
        self.running_var = torch.rand(c, device=dev) + 0.5
        self.eps = 1e-5


def nhwc(c, h=H, w=W):
    return torch.randn(1, c, h, w, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def run(name, fn, iters=3):
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print('%-10s %.3f ms' % (name, e0.elapsed_time(e1) / iters))


if which in ('expand', 'all'):
    x, w, bn, c = nhwc(64), torch.randn(384, 64, 1, 1, device=dev) * 0.1, BN(384), OF.PackedWeightCache()
    run('expand', lambda: OF.conv_bn_act_infer(x, w, 64, 384, 1, bn, B.ACT_RELU6, cache=c))
if which in ('project', 'all'):
    x, w, bn, c, r = nhwc(384), torch.randn(64, 384, 1, 1, device=dev) * 0.1, BN(64), OF.PackedWeightCache(), nhwc(64)
    run('project', lambda: OF.conv_bn_act_infer(x, w, 384, 64, 1, bn, B.ACT_NONE, residual=r, cache=c))
if which in ('dw7', 'all'):
    x, w7, bn = nhwc(384), torch.randn(384, 1, 7, 7, device=dev) * 0.1, BN(384)
    m75, m53 = torch.eye(25, device=dev), torch.eye(9, device=dev)
    run('dw7', lambda: OF.dw_bn_act_infer(x, w7, m75, m53, 7, True, bn, B.ACT_RELU6))
if which in ('dw3', 'all'):
    x, w7, bn = nhwc(192), torch.randn(384, 1, 7, 7, device=dev) * 0.1, BN(384)
    m75, m53 = torch.eye(25, device=dev), torch.eye(9, device=dev)
    run('dw3', lambda: OF.dw_bn_act_infer(x, w7, m75, m53, 3, True, bn, B.ACT_RELU6))
if which in ('ps', 'all'):
    x, w, bn, c = nhwc(64, 2 * H, 2 * W), torch.randn(256, 64, 5, 5, device=dev) * 0.02, BN(256), OF.PackedWeightCache()
    run('ps2x', lambda: OF.conv_bn_act_infer(x, w, 64, 256, 5, bn, B.ACT_NONE, B.STORE_PIXELSHUFFLE2, cache=c))
if which in ('out', 'all'):
    x, w, bn, c = nhwc(64, 4 * H, 4 * W), torch.randn(3, 64, 5, 5, device=dev) * 0.02, BN(3), OF.PackedWeightCache()
    run('out', lambda: OF.conv_bn_act_infer(x, w, 64, 3, 5, bn, B.ACT_NONE, cache=c, out_dtype=torch.float32, out_nchw=True))
