"""Golden of BatchNorm re-calibration: runs the UNMODIFIED reference's elastic_nn.utils.set_running_statistics
(elastic_nn/utils.py:16-66) on OFAMobileNetS4 (CPU) and stores the resulting running statistics.
    python tests/golden/make_golden_recal.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import ofa_sr_oracle as O  # noqa: E402
from ofa.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d  # noqa: E402
DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
from ofa.elastic_nn.networks.ofa_mbs4 import OFAMobileNetS4  # noqa: E402
from ofa.elastic_nn.utils import set_running_statistics  # noqa: E402

FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])
net = OFAMobileNetS4(pixelshuffle_depth_list=[1, 2], **FULL)
spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
net.load_state_dict(O.synth_state_dict(spec.param_shapes(), 95))
net.eval()
net.set_active_subnet(ks=5, e=4, d=3, pixel_d=2)
rs = np.random.RandomState(9)
loader = [{'image': torch.from_numpy(rs.rand(3, 3, 8, 12).astype(np.float32))},
          {'image': torch.from_numpy(rs.rand(2, 3, 8, 12).astype(np.float32))}]
set_running_statistics(net, loader)
out = {}
for k, v in net.state_dict().items():
    if k.endswith('running_mean') or k.endswith('running_var'):
        out[k] = v.numpy()
np.savez_compressed(os.path.join(HERE, 'reference_recal.npz'), **out)
print(len(out), 'buffers')
