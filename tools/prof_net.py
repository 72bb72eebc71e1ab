"""Per-kernel time table of one net forward (CUDA events around every library call): python tools/prof_net.py x4|s4 H W"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200'))
import numpy as np, torch
import ofa_b200
from ofa_b200 import functional as OF
from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
from ofa_b200.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
kind = sys.argv[1]; H, W = int(sys.argv[2]), int(sys.argv[3])
dev = torch.device('cuda:0')
DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])
net = (OFAMobileNetS4 if kind == 's4' else OFAMobileNetX4)(pixelshuffle_depth_list=[1, 2], **FULL).to(dev).eval()
net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
x = torch.rand(1, 3, H, W, device=dev)
with torch.no_grad():
    for _ in range(3):
        net(x)
    rec = []
    OF.set_profiler(rec)
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    OF.set_profiler(None)
agg = {}
for tag, flops, nbytes, e0, e1 in rec:
    a = agg.setdefault(tag, [0.0, 0, flops, nbytes]); a[0] += e0.elapsed_time(e1); a[1] += 1
tot = sum(a[0] for a in agg.values())
for tag, (ms, cnt, fl, nb) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%-36s x%-3d %.3f ms each  %5.1f%%  %6.0f GB/s %6.0f TFLOP/s' % (tag, cnt // 3, ms / cnt, 100 * ms / tot, nb / (ms / cnt) / 1e6, fl / (ms / cnt) / 1e9))
print('total %.3f ms per forward' % (tot / 3))
