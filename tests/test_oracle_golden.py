"""The oracle (oracle/ofa_sr_oracle.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only.  fp32 tolerance: 1e-5 relative to the output range (same
torch CPU kernels on both sides; differences are summation-order noise only)."""
import random

import numpy as np
import pytest
import torch

import ofa_sr_oracle as O

FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])


def _close(a, b, tol=2e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    scale = max(1.0, float(np.abs(b).max()))
    assert float(np.abs(a - b).max()) <= tol * scale, float(np.abs(a - b).max())


def _spec(kind, pd):
    return O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], pd)


def _apply(spec, request):
    if isinstance(request, str):
        random.seed(int(request.split(':')[1]))
        return spec.sample_active_subnet()
    spec.set_active_subnet(**request)
    return dict(request)


@pytest.mark.parametrize('name', ['s4_ps1', 's4_ps12', 'x4_ps12'])
def test_net_forward_and_bookkeeping(golden, name):
    arrays, book = golden
    meta = book[name]
    spec = _spec(meta['kind'], meta['pd'])
    shapes = spec.param_shapes()
    assert list(shapes.keys()) == meta['keys']
    sd = O.synth_state_dict(shapes, meta['wseed'])
    x = torch.from_numpy(arrays[name + '/x'])
    for i, sub in enumerate(meta['subnets']):
        setting = _apply(spec, sub['request'])
        assert setting == sub['setting']                       # incl. the mutated depth list (Q3)
        assert spec.runtime_depth == sub['runtime_depth']      # Q2
        active = [[spec.active_ks[b], spec.active_e[b]] for b in spec.mb_blocks]
        assert active == sub['active']                         # Q4
        with torch.no_grad():
            y = O.supernet_forward(x, sd, spec)
        assert list(y.shape) == sub['out_shape']               # Q1: 2-shuffle S4 always 4x
        _close(y.numpy(), arrays['%s/y%d' % (name, i)])


@pytest.mark.parametrize('key,kind,pd', [('sampling_s4_ps12', 's4', [1, 2]), ('sampling_x4_ps12', 'x4', [1, 2]),
                                         ('sampling_s4_ps1', 's4', [1])])
def test_sampling_stream_bit_exact(golden, key, kind, pd):
    _, book = golden
    spec = _spec(kind, pd)
    for row in book[key]:
        if 'constraint' in row:
            continue
        random.seed(row['seed'])
        assert spec.sample_active_subnet() == row['setting']
        assert spec.runtime_depth == row['runtime_depth']


def _block_sd():
    spec = _spec('s4', [1])
    blk = {k: v for k, v in spec.param_shapes().items() if k.startswith('blocks.0.mobile_inverted_conv.')}
    short = {k[len('blocks.0.mobile_inverted_conv.'):]: v for k, v in blk.items()}
    sd = O.synth_state_dict(short, 21)
    return {'blocks.0.mobile_inverted_conv.' + k: v for k, v in sd.items()}


def test_block_eval_and_filters(golden):
    arrays, _ = golden
    sd = _block_sd()
    x = torch.from_numpy(arrays['block/x'])
    p = 'blocks.0.mobile_inverted_conv.depth_conv.conv.'
    mats = {'7to5_matrix': sd[p + '7to5_matrix'], '5to3_matrix': sd[p + '5to3_matrix']}
    for ks in (3, 5, 7):
        f = O.active_filter(sd[p + 'conv.weight'], mats, [3, 5, 7], 384, ks)
        _close(f.numpy(), arrays['block/filter_k%d' % ks], 1e-6)
        for e in (3, 4, 6):
            with torch.no_grad():
                # the golden is the bare DynamicMBConvLayer (no residual): subtract x
                y = O.mbconv_block(x, sd, 'blocks.0.', ks, e, [3, 5, 7]) - x
            _close(y.numpy(), arrays['block/eval_k%d_e%d' % (ks, e)])


@pytest.mark.parametrize('ks,e', [(3, 4), (5, 6), (7, 3)])
def test_block_training_step(golden, ks, e):
    arrays, _ = golden
    tag = 'block/train_k%d_e%d/' % (ks, e)
    sd = _block_sd()
    pre = 'blocks.0.mobile_inverted_conv.'
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype == torch.float32 and 'running' not in k}
    x = torch.from_numpy(arrays['block/x']).requires_grad_(True)
    y = O.mbconv_block(x, sd, 'blocks.0.', ks, e, [3, 5, 7], training=True) - x
    loss = torch.nn.functional.mse_loss(y, torch.from_numpy(arrays[tag + 'target']))
    loss.backward()
    _close(y.detach().numpy(), arrays[tag + 'y'])
    assert abs(loss.item() - float(arrays[tag + 'loss'])) < 1e-5 * max(1.0, abs(loss.item()))
    # block(x) - x: the identity residual and the subtraction cancel, so x.grad is the bare layer's dx
    _close(x.grad.numpy(), arrays[tag + 'dx'], 1e-4)
    for k, v in params.items():
        ref = arrays[tag + 'grad/' + k[len(pre):]]
        if ref.size == 0:
            assert v.grad is None or float(v.grad.abs().max()) == 0.0, k   # e.g. 5to3 matrix when ks >= 5
        else:
            _close(v.grad.numpy(), ref, 1e-4)
    for name in ('depth_conv.bn.bn.running_mean', 'depth_conv.bn.bn.running_var',
                 'inverted_bottleneck.bn.bn.running_var', 'point_linear.bn.bn.running_mean'):
        _close(sd[pre + name].detach().numpy(), arrays[tag + 'buf/' + name], 1e-5)
    assert int(sd[pre + 'depth_conv.bn.bn.num_batches_tracked']) == int(arrays[tag + 'buf/depth_conv.bn.bn.num_batches_tracked'])


def test_pixel_reorders_are_inverse():
    x = torch.randn(2, 16, 6, 8)
    assert torch.equal(O.pixel_shuffle2(x), torch.nn.functional.pixel_shuffle(x, 2))
    assert torch.equal(O.pixel_unshuffle2(x), torch.nn.functional.pixel_unshuffle(x, 2))
    assert torch.equal(O.pixel_unshuffle2(O.pixel_shuffle2(x)), x)


def test_psnr_metric():
    a = torch.rand(3, 8, 8)
    ya = O.tensor_to_y_uint8(a)
    assert ya.dtype == np.uint8 and ya.shape == (8, 8)
    assert O.psnr_uint8(ya, ya) == float('inf')


def test_metric_matches_reference_functions():
    """oracle psnr_y_* vs psnr(rgb2y(tensor2img_np(.)), ...) of the unmodified reference (make_golden_metric.py),
    including the make_grid padding quirk for batches."""
    import json
    import math
    import os
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_metric.json')) as f:
        cases = json.load(f)
    for c in cases:
        rs = np.random.RandomState(c['seed'])
        a = (rs.rand(*c['shape']) * 1.2 - 0.1).astype(np.float32)
        b = (a + c['noise'] * rs.randn(*c['shape'])).astype(np.float32)
        n, _, h, w = c['shape']
        got = O.psnr_y_from_sse(O.psnr_y_sse(torch.from_numpy(a), torch.from_numpy(b)), n, h, w)
        if c['psnr'] is None:
            assert math.isinf(got)
        else:
            assert abs(got - c['psnr']) < 1e-12, (c, got)


def test_bn_recalibration_matches_reference():
    """oracle set_running_statistics vs the unmodified reference's (make_golden_recal.py): every running_mean /
    running_var of the S4 supernet after re-calibrating the (ks=5, e=4, d=3, pixel_d=2) subnet on two batches."""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_recal.npz'))
    FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
    sd = O.synth_state_dict(spec.param_shapes(), 95)
    spec.set_active_subnet(ks=5, e=4, d=3, pixel_d=2)
    rs = np.random.RandomState(9)
    batches = [torch.from_numpy(rs.rand(3, 3, 8, 12).astype(np.float32)), torch.from_numpy(rs.rand(2, 3, 8, 12).astype(np.float32))]
    O.set_running_statistics(sd, spec, batches)
    assert len(gold.files) == 108
    for k in gold.files:
        np.testing.assert_allclose(sd[k].numpy(), gold[k], rtol=2e-5, atol=1e-6, err_msg=k)
