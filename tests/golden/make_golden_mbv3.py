"""Golden vectors for the MobileNetV3 flavour of the elastic modules: runs the UNMODIFIED reference
modules on the CPU — DynamicMBConvLayer with stride 2 / squeeze-and-excite / h-swish, DynamicLinearLayer — forward and
backward (training mode, batch statistics), and stores parameters, inputs, outputs, gradients and the BN running
statistics after the step.
    python tests/golden/make_golden_mbv3.py         (build container only; writes reference_mbv3.npz)"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from ofa.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d  # noqa: E402
DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
from ofa.elastic_nn.modules.dynamic_layers import DynamicMBConvLayer, DynamicLinearLayer  # noqa: E402

out = {}


def randomize(mod, seed):
    rs = np.random.RandomState(seed)
    for k, v in mod.state_dict().items():
        if k.endswith('num_batches_tracked'):
            continue
        if k.endswith('running_var'):
            v.copy_(torch.from_numpy(rs.uniform(0.5, 1.5, tuple(v.shape)).astype(np.float32)))
        elif k.endswith('_matrix'):
            n = v.shape[0]
            v.copy_(torch.from_numpy((np.eye(n) + 0.05 * rs.randn(n, n)).astype(np.float32)))
        elif k.endswith('bn.weight'):
            v.copy_(torch.from_numpy(rs.uniform(0.5, 1.5, tuple(v.shape)).astype(np.float32)))
        elif v.dim() == 4:
            fan = v.shape[1] * v.shape[2] * v.shape[3]
            v.copy_(torch.from_numpy((rs.randn(*v.shape) * (1.5 / fan) ** 0.5).astype(np.float32)))
        elif v.dim() == 2:
            v.copy_(torch.from_numpy((rs.randn(*v.shape) * (1.0 / v.shape[1]) ** 0.5).astype(np.float32)))
        else:
            v.copy_(torch.from_numpy((0.2 * rs.randn(*v.shape)).astype(np.float32)))


CASES = [
    # name, in, out, stride, act, se, ks, e, shape
    ('se_s2_hswish', 24, 40, 2, 'h_swish', True, 5, 4, (3, 24, 13, 10)),
    ('se_s1_relu', 40, 40, 1, 'relu', True, 3, 3, (2, 40, 9, 7)),
    ('plain_s2_relu6', 16, 24, 2, 'relu6', False, 7, 6, (2, 16, 12, 12)),
    ('se_s1_hswish_k7', 32, 32, 1, 'h_swish', True, 7, 6, (2, 32, 8, 8)),
]
for name, cin, cout, stride, act, se, ks, e, shape in CASES:
    torch.manual_seed(0)
    m = DynamicMBConvLayer([cin], [cout], [3, 5, 7], [3, 4, 6], stride=stride, act_func=act, use_se=se)
    randomize(m, sum(ord(ch) for ch in name))
    for k, v in m.state_dict().items():
        out['%s/param/%s' % (name, k)] = v.numpy().copy()
    m.active_kernel_size, m.active_expand_ratio, m.active_out_channel = ks, e, cout
    rs = np.random.RandomState(7)
    x = torch.from_numpy(rs.randn(*shape).astype(np.float32)).requires_grad_(True)
    m.eval()
    with torch.no_grad():
        out[name + '/y_eval'] = m(x).numpy().copy()
        sub = m.get_active_subnet(cin)           # the reference's second statement of the active sub-network
        sub.eval()
        out[name + '/y_sub_eval'] = sub(x).numpy().copy()
    m.train()
    y = m(x)
    gy = torch.from_numpy(rs.randn(*y.shape).astype(np.float32))
    y.backward(gy)
    out[name + '/x'] = x.detach().numpy().copy()
    out[name + '/gy'] = gy.numpy().copy()
    out[name + '/y_train'] = y.detach().numpy().copy()
    out[name + '/dx'] = x.grad.numpy().copy()
    for k, p in m.named_parameters():
        if p.grad is not None:
            out['%s/grad/%s' % (name, k)] = p.grad.numpy().copy()
    for k, v in m.state_dict().items():
        if 'running_' in k:
            out['%s/after/%s' % (name, k)] = v.numpy().copy()

# DynamicLinearLayer
lin = DynamicLinearLayer([48, 64, 96], 10, bias=True, dropout_rate=0)
randomize(lin, 5)
for k, v in lin.state_dict().items():
    out['linear/param/%s' % k] = v.numpy().copy()
rs = np.random.RandomState(3)
for width in (48, 96):
    x = torch.from_numpy(rs.randn(5, width).astype(np.float32)).requires_grad_(True)
    lin.zero_grad()
    y = lin(x)
    gy = torch.from_numpy(rs.randn(*y.shape).astype(np.float32))
    y.backward(gy)
    out['linear/%d/x' % width], out['linear/%d/gy' % width] = x.detach().numpy().copy(), gy.numpy().copy()
    out['linear/%d/y' % width], out['linear/%d/dx' % width] = y.detach().numpy().copy(), x.grad.numpy().copy()
    out['linear/%d/dw' % width] = lin.linear.linear.weight.grad.numpy().copy()
    out['linear/%d/db' % width] = lin.linear.linear.bias.grad.numpy().copy()
np.savez_compressed(os.path.join(HERE, 'reference_mbv3.npz'), **out)
print('wrote', len(out), 'arrays')
