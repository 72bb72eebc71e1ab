"""Small-image / launch-bound cases: C1 (S4 smallest subnet, 2x, 1x3x256x256) and an X4 forward, eager vs captured
in a CUDA graph (torch.cuda.graph around net(x): the library launches on the capturing stream)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import ofa_b200, ofa_sr_oracle as O
from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
from ofa_b200.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
dev = torch.device('cuda:0')
DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])


def build(kind, pd):
    cls = OFAMobileNetS4 if kind == 's4' else OFAMobileNetX4
    net = cls(pixelshuffle_depth_list=list(pd), **{k: list(v) for k, v in FULL.items()})
    spec = O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], list(pd))
    net.load_state_dict(O.synth_state_dict(spec.param_shapes(), 5))
    return net.to(dev).eval()


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


cases = [('C1  S4 (ks3,e3,d2) 2x 256x256 -> 512x512', 's4', [1], dict(ks=3, e=3, d=2, pixel_d=1), (1, 3, 256, 256)),
         ('C1b S4 (ks3,e3,d2) 4x 256x256 -> 1024x1024', 's4', [1, 2], dict(ks=3, e=3, d=2, pixel_d=2), (1, 3, 256, 256)),
         ('C4  X4 max 4x down->up 1024x1024', 'x4', [1, 2], dict(ks=7, e=6, d=4, pixel_d=2), (1, 3, 1024, 1024)),
         ('C4  X4 (ks3,e3,d2) 2x down->up 1024x1024', 'x4', [1, 2], dict(ks=3, e=3, d=2, pixel_d=1), (1, 3, 1024, 1024))]
for name, kind, pd, sub, shape in cases:
    net = build(kind, pd)
    net.set_active_subnet(**sub)
    x = torch.rand(*shape, device=dev)
    with torch.no_grad():
        y = net(x)
        eager = timeit(lambda: net(x))
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                net(x)
        torch.cuda.current_stream().wait_stream(s)
        try:
            with torch.cuda.graph(g):
                yg = net(x)
            graphed = timeit(g.replay)
            ok = bool(torch.equal(yg, y))
            msg = 'graph %.3f ms (bit-identical: %s)' % (graphed, ok)
        except Exception as e:  # noqa: BLE001
            msg = 'graph capture failed: %r' % (e,)
    mpix = y.shape[0] * y.shape[2] * y.shape[3] / 1e6
    print('%-48s out %s  eager %.3f ms (%.0f Mpix/s)  %s' % (name, tuple(y.shape), eager, mpix / eager * 1e3, msg), flush=True)
# empty batch
net = build('s4', [1, 2]); net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
try:
    with torch.no_grad():
        y0 = net(torch.rand(0, 3, 16, 16, device=dev))
    print('empty batch ->', tuple(y0.shape))
except Exception as e:  # noqa: BLE001
    print('empty batch raised', repr(e))
