"""CPU-only checks of the host side: the drop-in module API, the bit-exact sub-network bookkeeping
(against fixtures of the unmodified reference) and the C-ABI library's exported surface.  No compute
entry point is exercised here — there is no GPU in this tier."""
import ctypes
import os
import random
import re

import pytest
import torch

import ofa_sr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])


def _net(kind, pd):
    from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
    from ofa_b200.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
    cls = OFAMobileNetS4 if kind == 's4' else OFAMobileNetX4
    return cls(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=list(pd))


@pytest.mark.parametrize('name', ['s4_ps1', 's4_ps12', 'x4_ps12'])
def test_state_dict_layout_matches_reference(golden, name):
    _, book = golden
    meta = book[name]
    net = _net(meta['kind'], meta['pd'])
    sd = net.state_dict()
    assert list(sd.keys()) == meta['keys']          # names AND registration order
    shapes = O.SuperNetSpec(meta['kind'], FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'],
                            meta['pd']).param_shapes()
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    assert all(v.dtype in (torch.float32, torch.int64) for v in sd.values())


@pytest.mark.parametrize('name', ['s4_ps1', 's4_ps12', 'x4_ps12'])
def test_set_active_subnet_bit_exact(golden, name):
    _, book = golden
    meta = book[name]
    net = _net(meta['kind'], meta['pd'])
    for sub in meta['subnets']:
        req = sub['request']
        if isinstance(req, str):
            random.seed(int(req.split(':')[1]))
            setting = net.sample_active_subnet()
        else:
            net.set_active_subnet(**req)
            setting = dict(req)
        assert setting == sub['setting']
        assert list(net.runtime_depth) == sub['runtime_depth']
        active = [[b.mobile_inverted_conv.active_kernel_size, b.mobile_inverted_conv.active_expand_ratio]
                  for b in net.blocks if hasattr(b, 'mobile_inverted_conv')]
        assert active == sub['active']


@pytest.mark.parametrize('key,kind,pd', [('sampling_s4_ps12', 's4', [1, 2]), ('sampling_x4_ps12', 'x4', [1, 2]),
                                         ('sampling_s4_ps1', 's4', [1])])
def test_sampling_stream_and_constraints(golden, key, kind, pd):
    _, book = golden
    net = _net(kind, pd)
    for row in book[key]:
        if 'constraint' in row:
            for ctype, lst in row['constraint'].items():
                net.set_constraint(lst, ctype)
        random.seed(row['seed'])
        assert net.sample_active_subnet() == row['setting']
        assert list(net.runtime_depth) == row['runtime_depth']
    net.clear_constraint()
    with pytest.raises(NotImplementedError):
        net.set_constraint([1], 'bogus')


def test_quirks_documented_in_survey():
    s4 = _net('s4', [1, 2])
    s4.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
    assert s4.runtime_depth == [4, 4, 4, 2, 2]            # Q2: 14 MB blocks, not 16
    d = [3, 3, 3, 4]
    s4.set_active_subnet(ks=3, e=3, d=d, pixel_d=1)
    assert d == [3, 3, 3, 1, 4]                           # Q3: caller's list mutated
    s4_1 = _net('s4', [1])
    s4_1.set_active_subnet(ks=3, e=3, d=2, pixel_d=1)
    assert s4_1.runtime_depth == [2, 2, 2, 1, 1]          # Q5 config C1
    assert s4_1.blocks[15].mobile_inverted_conv.active_kernel_size == 7   # Q4: last MB block untouched
    x4 = _net('x4', [1, 2])
    x4.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
    assert x4.runtime_depth[:4] == [2, 4, 4, 4]
    with pytest.raises(AttributeError):
        s4.get_active_subnet()                            # Q7


def test_module_api_surface():
    from ofa_b200.elastic_nn.modules import (DynamicMBConvLayer, DynamicSeparableConv2d, DynamicPointConv2d,
                                             DynamicBatchNorm2d, DynamicConvLayer)
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
    layer = DynamicMBConvLayer([64], [64], [3, 5, 7], [3, 4, 6])
    assert isinstance(layer.inverted_bottleneck.conv, DynamicPointConv2d)
    assert isinstance(layer.depth_conv.conv, DynamicSeparableConv2d)
    assert isinstance(layer.point_linear.bn, DynamicBatchNorm2d)
    assert layer.depth_conv.conv._ks_set == [3, 5, 7]
    assert {n for n, _ in layer.depth_conv.conv.named_parameters()} == {'conv.weight', '5to3_matrix', '7to5_matrix'}
    assert layer.active_kernel_size == 7 and layer.active_expand_ratio == 6 and layer.active_out_channel == 64
    assert layer.module_str == '(O64, E6.0, K7)'
    cfg = dict(layer.config)
    assert cfg.pop('name') == 'DynamicMBConvLayer'      # callers pop the name, as set_layer_from_config does
    assert DynamicMBConvLayer.build_from_config(cfg).config == layer.config
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = None
    plain = DynamicSeparableConv2d(8, [3, 5, 7])
    assert [n for n, _ in plain.named_parameters()] == ['conv.weight']     # matrices only when the mode is set
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
    assert DynamicBatchNorm2d.SET_RUNNING_STATISTICS is False
    assert DynamicConvLayer([3], [16], 3).module_str == 'DyConv(O16, K3, S1)'
    # the MobileNetV3 flavour: same class names, children and state_dict keys as the reference
    from ofa_b200.elastic_nn.modules import DynamicSE, DynamicLinear, DynamicLinearLayer
    from ofa_b200.layers import LinearLayer, MBInvertedConvLayer, set_layer_from_config
    from ofa_b200.utils import SEModule, make_divisible
    se = DynamicSE(96)
    assert isinstance(se, SEModule) and se.reduction == 4
    assert set(se.state_dict()) == {'fc.reduce.weight', 'fc.reduce.bias', 'fc.expand.weight', 'fc.expand.bias'}
    assert tuple(se.fc.reduce.weight.shape) == (make_divisible(96 // 4, 8), 96, 1, 1)
    v3 = DynamicMBConvLayer([24], [40], [3, 5, 7], [3, 4, 6], stride=2, act_func='h_swish', use_se=True)
    assert isinstance(v3.depth_conv.se, DynamicSE) and v3.depth_conv.conv.stride == 2
    assert list(v3.depth_conv._modules) == ['conv', 'bn', 'act', 'se']
    lin = DynamicLinearLayer([48, 96], 10, bias=True, dropout_rate=0.1)
    assert isinstance(lin.linear, DynamicLinear) and lin.linear.active_out_features == 10 and lin.dropout is not None
    static = MBInvertedConvLayer(16, 24, 5, stride=2, expand_ratio=4, act_func='h_swish', use_se=True)
    assert static.module_str == 'SE_5x5_MBConv4_H_SWISH_O24'
    cfg = LinearLayer(32, 10).config
    assert isinstance(set_layer_from_config(cfg), LinearLayer)


def test_load_weights_from_net_key_mapping():
    s4 = _net('s4', [1, 2])
    sd = s4.state_dict()
    wrapped = {'module.' + k: v.clone() + (1 if v.dtype == torch.float32 else 0) for k, v in sd.items()}
    s4.load_weights_from_net(wrapped)
    k = 'blocks.3.mobile_inverted_conv.point_linear.conv.conv.weight'
    assert torch.equal(s4.state_dict()[k], wrapped['module.' + k])


def test_library_exports_every_declared_symbol():
    from ofa_b200 import backend as B
    header = open(os.path.join(ROOT, 'include', 'ofa_sr_b200.h')).read()
    declared = set(re.findall(r'\b(ofa_[a-z0-9_]+)\s*\(', header))
    assert declared == set(B.SYMBOLS), declared ^ set(B.SYMBOLS)
    assert os.path.exists(B.LIB_PATH), 'build the library first (__graft_entry__.build())'
    lib = ctypes.CDLL(B.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert B.lib().ofa_version() >= 100


def test_no_cpu_fallback():
    """A CPU tensor must be refused by the binding, and the library itself reports OFA_ERR_CUDA when
    no device exists."""
    from ofa_b200 import backend as B, functional as OF
    with pytest.raises(RuntimeError):
        B.t4(torch.zeros(1, 1, 1, 1))
    if not torch.cuda.is_available():
        sm = ctypes.c_int32()
        rc = B.lib().ofa_device_info(ctypes.byref(sm), None, None)
        assert rc == 2 and b'no CPU fallback' in B.lib().ofa_last_error()
        net = _net('s4', [1]).eval()
        with torch.no_grad(), pytest.raises(RuntimeError):
            net(torch.rand(1, 3, 8, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'ofa-for-super-resolution_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'ofa_sr_oracle' not in text and 'import oracle' not in text, os.path.join(dirpath, f)


def test_planar_dispatch_rule():
    """IMPL_AUTO's frame / patch rule (functional.planar_preferred, mirrored by mbconv_planar_preferred in the library):
    planes of >= 2304 pixels (48 x 48) that fill >= 25 % of their 128 (64-row tail) x 112 depthwise tiles take the planar
    tensor-core path; batches of small patches take the NHWC kernels."""
    from ofa_b200 import functional as OF
    mk = lambda h, w: torch.empty(1, 64, h, w, dtype=torch.float16)
    for h, w, want in [(24, 24, False), (48, 48, True), (40, 56, False), (32, 64, False), (8, 2048, False), (64, 128, True), (64, 64, True),
                       (96, 96, True), (96, 120, True), (256, 256, True), (540, 960, True), (1080, 1920, True),
                       (129, 56, True), (129, 24, False), (129, 64, True), (130, 232, True)]:
        assert OF.planar_preferred(mk(h, w)) is want, (h, w)
    assert OF.planar_supported(mk(24, 24), 64, 384, 64) and not OF.planar_supported(mk(24, 25), 64, 384, 64)
    assert not OF.planar_supported(mk(24, 24).float(), 64, 384, 64) and not OF.planar_supported(mk(24, 24), 64, 400, 64)
