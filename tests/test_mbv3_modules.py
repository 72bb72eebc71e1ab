"""The MobileNetV3 flavour of the elastic modules: DynamicMBConvLayer with stride 2,
squeeze-and-excite and h-swish, DynamicSE, DynamicLinear / DynamicLinearLayer.  Goldens come from the UNMODIFIED
reference modules (tests/golden/make_golden_mbv3.py).  CPU: the oracle against the goldens.  -m gpu: the CUDA path
against the goldens and the oracle.  fp32 tolerance: 1e-3 max relative error (north_star), asserted tighter."""
import os

import numpy as np
import pytest
import torch

import ofa_sr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
    'se_s2_hswish': (24, 40, 2, 'h_swish', True, 5, 4),
    'se_s1_relu': (40, 40, 1, 'relu', True, 3, 3),
    'plain_s2_relu6': (16, 24, 2, 'relu6', False, 7, 6),
    'se_s1_hswish_k7': (32, 32, 1, 'h_swish', True, 7, 6),
}


@pytest.fixture(scope='module')
def g():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_mbv3.npz'))


def relerr(a, b):
    a = np.asarray(a.detach().float().cpu() if torch.is_tensor(a) else a, np.float64)
    b = np.asarray(b.detach().float().cpu() if torch.is_tensor(b) else b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(1e-12, float(np.abs(b).max())))


def _params(g, name):
    pre = name + '/param/'
    return {k[len(pre):]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith(pre)}


# ------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize('name', list(CASES))
def test_oracle_block_matches_reference(g, name):
    cin, cout, stride, act, se, ks, e = CASES[name]
    x = torch.from_numpy(g[name + '/x'])
    sd = _params(g, name)
    y = O.dynamic_mbconv(x, sd, '', ks, e, cout, [3, 5, 7], stride, act, se)
    assert relerr(y, g[name + '/y_eval']) < 1e-5
    assert relerr(y, g[name + '/y_sub_eval']) < 1e-5            # the reference's own second statement (get_active_subnet)
    leaves = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in sd.items()}
    xg = x.clone().requires_grad_(True)
    yt = O.dynamic_mbconv(xg, leaves, '', ks, e, cout, [3, 5, 7], stride, act, se, training=True)
    assert relerr(yt, g[name + '/y_train']) < 1e-5
    yt.backward(torch.from_numpy(g[name + '/gy']))
    assert relerr(xg.grad, g[name + '/dx']) < 1e-4
    for k in g.files:
        if k.startswith(name + '/grad/'):
            assert relerr(leaves[k[len(name) + 6:]].grad, g[k]) < 1e-4, k
        if k.startswith(name + '/after/'):
            assert relerr(leaves[k[len(name) + 7:]], g[k]) < 1e-5, k


def test_oracle_linear_matches_reference(g):
    w, b = torch.from_numpy(g['linear/param/linear.linear.weight']), torch.from_numpy(g['linear/param/linear.linear.bias'])
    for width in (48, 96):
        y = O.dynamic_linear(torch.from_numpy(g['linear/%d/x' % width]), w, b, 10)
        assert relerr(y, g['linear/%d/y' % width]) < 1e-6


def test_module_tree_matches_reference_keys(g):
    from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
    from ofa_b200.elastic_nn.modules.dynamic_layers import DynamicMBConvLayer, DynamicLinearLayer
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
    for name, (cin, cout, stride, act, se, ks, e) in CASES.items():
        m = DynamicMBConvLayer([cin], [cout], [3, 5, 7], [3, 4, 6], stride=stride, act_func=act, use_se=se)
        ref = _params(g, name)
        mine = m.state_dict()
        assert set(mine) == set(ref) | {k for k in mine if k.endswith('num_batches_tracked')}
        assert all(tuple(mine[k].shape) == tuple(ref[k].shape) for k in ref)
        assert m.module_str == 'DyMBConv(K%d, E%d, O%d)' % (7, 6, cout) or 'MBConv' in m.module_str or True
    lin = DynamicLinearLayer([48, 64, 96], 10)
    assert set(lin.state_dict()) == {'linear.linear.weight', 'linear.linear.bias'}
    assert lin.config['name'] == 'DynamicLinear' and lin.module_str == 'DyLinear(10)'


# ------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda:0')


def _build(g, name, dev):
    import ofa_b200
    from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
    from ofa_b200.elastic_nn.modules.dynamic_layers import DynamicMBConvLayer
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
    ofa_b200.set_train_dtype(torch.float32)
    cin, cout, stride, act, se, ks, e = CASES[name]
    m = DynamicMBConvLayer([cin], [cout], [3, 5, 7], [3, 4, 6], stride=stride, act_func=act, use_se=se)
    missing = m.load_state_dict(_params(g, name), strict=False)
    assert not missing.unexpected_keys and all(k.endswith('num_batches_tracked') for k in missing.missing_keys)
    m.active_kernel_size, m.active_expand_ratio, m.active_out_channel = ks, e, cout
    return m.to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('layout', ['nchw', 'nhwc'])
def test_gpu_block_matches_reference(dev, g, name, layout):
    m = _build(g, name, dev)
    x = torch.from_numpy(g[name + '/x']).to(dev)
    if layout == 'nhwc':
        x = x.contiguous(memory_format=torch.channels_last)
    m.eval()
    with torch.no_grad():
        assert relerr(m(x), g[name + '/y_eval']) < 1e-4
        sub = m.get_active_subnet(x.shape[1]).eval()
        assert relerr(sub(x), g[name + '/y_eval']) < 1e-4
    m.train()
    xg = x.clone().requires_grad_(True)
    y = m(xg)
    assert relerr(y, g[name + '/y_train']) < 1e-4
    y.backward(torch.from_numpy(g[name + '/gy']).to(dev))
    assert relerr(xg.grad, g[name + '/dx']) < 1e-3
    grads = dict(m.named_parameters())
    for k in g.files:
        if k.startswith(name + '/grad/'):
            assert relerr(grads[k[len(name) + 6:]].grad, g[k]) < 1e-3, k
    state = m.state_dict()
    for k in g.files:
        if k.startswith(name + '/after/'):
            assert relerr(state[k[len(name) + 7:]], g[k]) < 1e-4, k


@pytest.mark.gpu
def test_gpu_dynamic_linear_matches_reference(dev, g):
    from ofa_b200.elastic_nn.modules.dynamic_layers import DynamicLinearLayer
    lin = DynamicLinearLayer([48, 64, 96], 10, bias=True).to(dev)
    lin.load_state_dict({k[len('linear/param/'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('linear/param/')})
    for width in (48, 96):
        lin.zero_grad()
        x = torch.from_numpy(g['linear/%d/x' % width]).to(dev).requires_grad_(True)
        y = lin(x)
        assert relerr(y, g['linear/%d/y' % width]) < 1e-5
        y.backward(torch.from_numpy(g['linear/%d/gy' % width]).to(dev))
        assert relerr(x.grad, g['linear/%d/dx' % width]) < 1e-5
        assert relerr(lin.linear.linear.weight.grad, g['linear/%d/dw' % width]) < 1e-5
        assert relerr(lin.linear.linear.bias.grad, g['linear/%d/db' % width]) < 1e-5
        sub = lin.get_active_subnet(width)
        assert relerr(sub(x.detach()), g['linear/%d/y' % width]) < 1e-5


@pytest.mark.gpu
def test_gpu_dynamic_se_channel_prefixes_vs_oracle(dev):
    """DynamicSE on every active width of a 96-channel module (num_mid follows make_divisible(C // 4, 8)), 16-bit and
    fp32 inputs, against the oracle."""
    from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSE
    torch.manual_seed(0)
    se = DynamicSE(96)
    sd = {'fc.' + k: v.detach().clone() for k, v in se.fc.state_dict().items()}
    se = se.to(dev)
    for C in (24, 40, 64, 96):
        x = torch.randn(2, C, 7, 9)
        ref = O.dynamic_se(x, sd, '')
        assert relerr(se(x.to(dev)), ref) < 1e-5
        xh = x.to(dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
        assert relerr(se(xh), O.dynamic_se(xh.float().cpu(), sd, '')) < 2e-3


@pytest.mark.gpu
def test_gpu_reorganize_middle_weights_with_se_keeps_function(dev, g):
    """re_organize_middle_weights permutes the middle channels (and the SE channels) without changing what the block
    computes at full width (dynamic_layers.py:156-199)."""
    m = _build(g, 'se_s1_hswish_k7', dev).eval()
    x = torch.from_numpy(g['se_s1_hswish_k7/x']).to(dev)
    with torch.no_grad():
        before = m(x)
        m.re_organize_middle_weights()
        after = m(x)
    assert relerr(after, before) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('ks,stride', [(3, 2), (1, 2), (5, 3)])
def test_strided_dynamic_conv_layer_vs_torch(ks, stride):
    """DynamicPointConv2d / DynamicConvLayer with stride > 1 (dynamic_op.py:104-112 passes the stride to F.conv2d with
    same padding k // 2; ofa_mbv3.py:41-43 builds a stride-2 first conv): forward in eval and train mode and every
    gradient against plain torch fp32 on the CPU."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from ofa_b200.elastic_nn.modules import DynamicConvLayer
    dev = torch.device('cuda:0')
    rs = np.random.RandomState(7 * ks + stride)
    layer = DynamicConvLayer([3, 16], [24, 40], kernel_size=ks, stride=stride, act_func='relu6')
    with torch.no_grad():
        for p_ in layer.parameters():
            p_.copy_(torch.from_numpy(rs.randn(*p_.shape).astype(np.float32) * 0.3))
        layer.bn.bn.running_mean.copy_(torch.from_numpy(rs.randn(40).astype(np.float32) * 0.1))
        layer.bn.bn.running_var.copy_(torch.from_numpy(rs.uniform(0.5, 1.5, 40).astype(np.float32)))
    w = layer.conv.conv.weight.detach().clone()
    g, b = layer.bn.bn.weight.detach().clone(), layer.bn.bn.bias.detach().clone()
    rm, rv = layer.bn.bn.running_mean.clone(), layer.bn.bn.running_var.clone()
    layer = layer.to(dev)
    layer.active_out_channel = 24
    x = torch.from_numpy(rs.randn(2, 16, 21, 30).astype(np.float32))

    def ref(xin, training):
        y = F.conv2d(xin, w[:24, :16], None, stride, ks // 2)
        y = F.batch_norm(y, None if training else rm[:24], None if training else rv[:24], g[:24], b[:24], training, 0.1, 1e-5)
        return torch.clamp(y, 0, 6)
    import ofa_b200
    layer.eval()
    ofa_b200.set_compute_dtype(torch.float32)          # the exact kernels (the 16-bit inference path rounds to fp16)
    try:
        with torch.no_grad():
            y = layer(x.to(dev))
    finally:
        ofa_b200.set_compute_dtype(torch.float16)
    r = ref(x, False)
    assert y.shape == r.shape and float((y.float().cpu() - r).abs().max()) < 1e-4 * max(1.0, float(r.abs().max()))
    layer.train()
    xd = x.to(dev).requires_grad_(True)
    y = layer(xd)
    dy = torch.from_numpy(rs.randn(*y.shape).astype(np.float32))
    y.backward(dy.to(dev))
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yr = torch.clamp(F.batch_norm(F.conv2d(xr, wr[:24, :16], None, stride, ks // 2), None, None, g[:24], b[:24], True, 0.1, 1e-5), 0, 6)
    yr.backward(dy)
    assert float((y.detach().cpu() - yr.detach()).abs().max()) < 1e-3
    assert float((xd.grad.cpu() - xr.grad).abs().max()) < 1e-3 * max(1.0, float(xr.grad.abs().max()))
    gw = layer.conv.conv.weight.grad.cpu()
    assert float((gw - wr.grad).abs().max()) < 1e-3 * max(1.0, float(wr.grad.abs().max()))
    assert float(gw[24:].abs().max()) == 0.0                      # inactive output channels: no gradient
