"""X4 teacher (max sub-network, eval, no_grad) on a batch of 64 96x96 HR patches -- the frozen-teacher forwards of the C4
step: planes of 48x48 (2x) and 24x24 (4x) pixels.  IMPL_AUTO (planar_preferred rule) against the planar kernels forced
(IMPL_FAST) and the NHWC kernels forced (OFA_PLANAR_MIN_AREA=huge)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch
import ofa_b200
import ofa_sr_oracle as O
from ofa_b200 import backend as B
from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
from ofa_b200.elastic_nn.networks import OFAMobileNetX4
dev = torch.device('cuda', 0)
DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
cfg = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=[1, 2])
net = OFAMobileNetX4(**{k: list(v) for k, v in cfg.items()})
spec = O.SuperNetSpec('x4', cfg['ks_list'], cfg['expand_ratio_list'], cfg['depth_list'], [1, 2])
net.load_state_dict(O.synth_state_dict(spec.param_shapes(), 7))
net = net.to(dev).eval()
x = torch.rand(64, 3, 96, 96, device=dev)
for pd in (1, 2):
    net.set_active_subnet(ks=7, e=6, d=4, pixel_d=pd)
    for name, impl in (('auto', B.IMPL_AUTO), ('planar forced', B.IMPL_FAST)):
        ofa_b200.set_impl(impl)
        with torch.no_grad():
            for _ in range(3):
                y = net(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                y = net(x)
            e1.record()
            torch.cuda.synchronize()
        print('X4 teacher pixel_d=%d (%dx%d planes) %-14s %.3f ms per forward' % (pd, 96 >> pd, 96 >> pd, name, e0.elapsed_time(e1) / 10))
ofa_b200.set_impl(B.IMPL_AUTO)
