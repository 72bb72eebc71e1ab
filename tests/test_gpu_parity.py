"""Parity of the CUDA path (called through the C ABI of libofa_sr_b200.so) with the CPU oracle and
with the fixtures produced by the unmodified reference.  Needs a B200: `pytest -m gpu`.

Tolerances (BASELINE.json north_star):
  fp32 path : max |err| <= 1e-3 * max|ref|              (we assert 1e-4 where the math is short)
  bf16 path : per-op   max |err| <= 2^-7 * max|ref| (one bf16 rounding of the output + bf16 inputs)
  fp16 path : per-op   max |err| <= 2^-9 * max|ref| (fp16 operands and output)
  16-bit    : per-net  PSNR(ours, target) within 0.01 dB of PSNR(reference, target), reference metric —
              asserted for fp16 storage; bf16 storage is held to 0.03 dB (its operand rounding alone
              measures 0.012-0.019 dB on these 14/28-block random nets)
  subnet selection / channel indexing: bit-exact
"""
import ctypes
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import ofa_sr_oracle as O

FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda:0')


@pytest.fixture(autouse=True)
def _reset_policy():
    import ofa_b200
    from ofa_b200 import backend as B
    from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
    ofa_b200.set_compute_dtype(torch.bfloat16)
    ofa_b200.set_impl(B.IMPL_AUTO)
    yield
    ofa_b200.set_compute_dtype(torch.float16)
    ofa_b200.set_impl(B.IMPL_AUTO)


def relerr(a, b):
    a = a.detach().float().cpu().double()
    b = b.detach().float().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / max(1e-12, float(b.abs().max())))


def bf16r(t):
    return t.to(torch.bfloat16).float()


def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy((np.random.RandomState(seed).randn(*shape) * scale).astype(np.float32))


def bn_params(C, seed, dev):
    rs = np.random.RandomState(seed)
    mk = lambda a: torch.from_numpy(a.astype(np.float32)).to(dev)
    return dict(gamma=mk(rs.uniform(0.5, 1.5, C)), beta=mk(0.1 * rs.randn(C)), mean=mk(0.1 * rs.randn(C)),
                var=mk(rs.uniform(0.5, 1.5, C)))


def ref_bn(y, bn, C, eps=1e-5):
    g, b, m, v = (bn[k][:C].cpu() for k in ('gamma', 'beta', 'mean', 'var'))
    return (y - m.view(1, -1, 1, 1)) / torch.sqrt(v.view(1, -1, 1, 1) + eps) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def dw_weights(seed, dev, C=384):
    rs = np.random.RandomState(seed)
    w7 = torch.from_numpy((rs.randn(C, 1, 7, 7) * 0.2).astype(np.float32))
    m75 = torch.from_numpy((np.eye(25) + 0.05 * rs.randn(25, 25)).astype(np.float32))
    m53 = torch.from_numpy((np.eye(9) + 0.05 * rs.randn(9, 9)).astype(np.float32))
    return w7, m75, m53


# =================================================================================================
# (a1) active filter
# =================================================================================================
@pytest.mark.parametrize('ks', [3, 5, 7])
@pytest.mark.parametrize('transform', [True, False])
def test_active_filter(dev, ks, transform):
    from ofa_b200 import functional as OF
    w7, m75, m53 = dw_weights(1, dev)
    ref = O.active_filter(w7, {'7to5_matrix': m75, '5to3_matrix': m53}, [3, 5, 7], 200, ks, transform)
    got = OF.dw_active_filter(w7.to(dev), m75.to(dev), m53.to(dev), transform, ks, 200)
    assert relerr(got, ref) < 1e-5


# =================================================================================================
# (a2) depthwise forward — CUDA-core path, every layout / dtype; TMA path, bf16 NHWC
# =================================================================================================
@pytest.mark.parametrize('ks', [3, 5, 7])
@pytest.mark.parametrize('layout', ['nchw', 'nhwc'])
@pytest.mark.parametrize('C', [40, 192])
def test_dw_simt_fp32(dev, ks, layout, C):
    from ofa_b200 import functional as OF, backend as B
    import ofa_b200
    ofa_b200.set_impl(B.IMPL_SIMT)
    w7, m75, m53 = dw_weights(2, dev)
    x = rnd(2, C, 13, 9, seed=3)
    filt = O.active_filter(w7, {'7to5_matrix': m75, '5to3_matrix': m53}, [3, 5, 7], C, ks)
    ref = O.dw_conv(x, filt)
    xd = x.to(dev)
    if layout == 'nhwc':
        xd = xd.contiguous(memory_format=torch.channels_last)
    got = OF.dw_conv(xd, w7.to(dev), m75.to(dev), m53.to(dev), ks, True)
    assert relerr(got, ref) < 1e-5
    # fused BN + ReLU6 epilogue
    bn = bn_params(384, 4, dev)

    class _BN:  # duck-typed nn.BatchNorm2d for the fused call
        weight, bias, running_mean, running_var, eps = bn['gamma'], bn['beta'], bn['mean'], bn['var'], 1e-5
    got2 = OF.dw_bn_act_infer(xd, w7.to(dev), m75.to(dev), m53.to(dev), ks, True, _BN, B.ACT_RELU6)
    ref2 = torch.clamp(ref_bn(ref, bn, C), 0, 6)
    assert relerr(got2, ref2) < 1e-5


def test_fp16_net_on_ragged_width_uses_nhwc_kernels(dev):
    """W % 8 != 0 takes the three NHWC kernels instead of the planar path (TMA needs 16-byte row pitches): the
    fp16 default must stay on the fast depthwise kernel there and agree with the oracle."""
    import ofa_b200
    ofa_b200.set_compute_dtype(torch.float16)
    net = _build_net('s4', [1, 2], 91, dev)
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
    sd = O.synth_state_dict(spec.param_shapes(), 91)
    x = torch.from_numpy(np.random.RandomState(2).rand(1, 3, 21, 45).astype(np.float32))
    for sub in (dict(ks=7, e=6, d=4, pixel_d=2), dict(ks=3, e=3, d=2, pixel_d=1)):
        net.set_active_subnet(**sub)
        spec.set_active_subnet(**sub)
        with torch.no_grad():
            y = net(x.to(dev))
        assert relerr(y, O.supernet_forward(x, sd, spec)) < 1e-2


@pytest.mark.parametrize('ks', [3, 5, 7])
@pytest.mark.parametrize('shape', [(1, 64, 8, 32), (2, 192, 19, 45), (1, 384, 40, 70), (2, 128, 24, 48), (3, 64, 12, 24)])
def test_dw_fast_bf16(dev, ks, shape):
    from ofa_b200 import functional as OF, backend as B
    import ofa_b200
    ofa_b200.set_impl(B.IMPL_FAST)
    n, C, H, W = shape
    w7, m75, m53 = dw_weights(5, dev)
    x = bf16r(rnd(n, C, H, W, seed=6))
    filt = O.active_filter(w7, {'7to5_matrix': m75, '5to3_matrix': m53}, [3, 5, 7], C, ks)
    bn = bn_params(384, 7, dev)
    ref = torch.clamp(ref_bn(O.dw_conv(x, filt), bn, C), 0, 6)

    class _BN:
        weight, bias, running_mean, running_var, eps = bn['gamma'], bn['beta'], bn['mean'], bn['var'], 1e-5
    xd = x.to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    got = OF.dw_bn_act_infer(xd, w7.to(dev), m75.to(dev), m53.to(dev), ks, True, _BN, B.ACT_RELU6)
    assert got.dtype == torch.bfloat16
    assert relerr(got, ref) < 2 ** -7
    # and the two implementations agree with each other to one bf16 ulp of the output range
    ofa_b200.set_impl(B.IMPL_SIMT)
    got_s = OF.dw_bn_act_infer(xd, w7.to(dev), m75.to(dev), m53.to(dev), ks, True, _BN, B.ACT_RELU6)
    assert relerr(got, got_s) < 2 ** -7


@pytest.mark.parametrize('ks', [3, 5, 7])
@pytest.mark.parametrize('shape', [(2, 40, 13, 9), (3, 192, 24, 24), (1, 200, 5, 2), (2, 384, 7, 31)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_dw_backward_nhwc(dev, ks, shape, dtype):
    """(a14) depthwise backward on dense NHWC tensors (the row-walk filter-gradient kernel, the data gradient and
    the chain rule through the 7->5->3 transforms) against autograd on the oracle's statement of the forward.
    Ragged cases: C not a multiple of 64, W smaller than the filter, rows shared between row groups."""
    from ofa_b200 import functional as OF, backend as B
    import ofa_b200
    ofa_b200.set_impl(B.IMPL_AUTO)
    n, C, H, W = shape
    w7, m75, m53 = dw_weights(11, dev)
    x = rnd(n, C, H, W, seed=12)
    gy = rnd(n, C, H, W, seed=13)
    if dtype == torch.bfloat16:
        x, gy = bf16r(x), bf16r(gy)
    leaves = [t.clone().requires_grad_(True) for t in (x, w7, m75, m53)]
    filt = O.active_filter(leaves[1], {'7to5_matrix': leaves[2], '5to3_matrix': leaves[3]}, [3, 5, 7], C, ks)
    O.dw_conv(leaves[0], filt).backward(gy)
    xd = x.to(dev).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wd = [t.to(dev).requires_grad_(True) for t in (w7, m75, m53)]
    y = OF.dw_conv(xd, wd[0], wd[1], wd[2], ks, True)
    y.backward(gy.to(dev).to(dtype).contiguous(memory_format=torch.channels_last))
    tol = 1e-4 if dtype == torch.float32 else 2 ** -7
    assert relerr(xd.grad, leaves[0].grad) < tol
    assert relerr(wd[0].grad, leaves[1].grad) < 1e-4          # fp32 accumulation of exactly representable products
    if ks < 7:
        assert relerr(wd[1].grad, leaves[2].grad) < 1e-4
    else:
        assert wd[1].grad is None and leaves[2].grad is None
    if ks == 3:
        assert relerr(wd[2].grad, leaves[3].grad) < 1e-4
    else:
        assert wd[2].grad is None and leaves[3].grad is None


# =================================================================================================
# (a3, a4, a9-a11) dense conv — CUDA-core path and tcgen05 path
# =================================================================================================
def _conv_case(dev, ks, cin, cout, cmax_in, cmax_out, store, with_res, shape, dtype, impl, seed=0):
    from ofa_b200 import functional as OF, backend as B
    import ofa_b200
    ofa_b200.set_impl(impl)
    n, H, W = shape
    w = rnd(cmax_out, cmax_in, ks, ks, seed=seed, scale=(2.0 / (ks * ks * cout)) ** 0.5)
    x = rnd(n, cin, H, W, seed=seed + 1)
    if dtype == torch.bfloat16:
        x, wq = bf16r(x), bf16r(w)
    else:
        wq = w
    bn = bn_params(cmax_out, seed + 2, dev)
    ref = ref_bn(O.sliced_conv(x, wq, cout), bn, cout)
    act = B.ACT_RELU6 if store == B.STORE_PLAIN else B.ACT_NONE
    if act == B.ACT_RELU6:
        ref = torch.clamp(ref, 0, 6)
    if store == B.STORE_PIXELSHUFFLE2:
        ref = O.pixel_shuffle2(ref)
    elif store == B.STORE_PIXELUNSHUFFLE2:
        ref = O.pixel_unshuffle2(ref)
    res = None
    if with_res:
        res = rnd(*ref.shape, seed=seed + 3)
        if dtype == torch.bfloat16:
            res = bf16r(res)
        ref = ref + res
        res = res.to(dev).to(dtype).contiguous(memory_format=torch.channels_last)

    class _BN:
        weight, bias, running_mean, running_var, eps = bn['gamma'], bn['beta'], bn['mean'], bn['var'], 1e-5
    xd = x.to(dev).to(dtype).contiguous(memory_format=torch.channels_last)
    cache = OF.PackedWeightCache()
    got = OF.conv_bn_act_infer(xd, w.to(dev), cin, cout, ks, _BN, act, store, res, cache, out_dtype=dtype)
    return got, ref


@pytest.mark.parametrize('ks,cin,cout', [(1, 64, 192), (1, 200, 64), (3, 3, 16), (5, 64, 3), (5, 16, 40), (7, 8, 8)])
@pytest.mark.parametrize('store', [0, 1, 2])
def test_conv_simt_fp32(dev, ks, cin, cout, store):
    from ofa_b200 import backend as B
    if store == B.STORE_PIXELSHUFFLE2 and cout % 4:
        pytest.skip('pixelshuffle needs cout % 4 == 0')
    got, ref = _conv_case(dev, ks, cin, cout, cin + 8, cout + 8, store, store == 0, (2, 10, 14), torch.float32, B.IMPL_SIMT)
    assert relerr(got, ref) < 1e-4


@pytest.mark.parametrize('ks,cin,cout,store,res', [
    (1, 64, 384, 0, False), (1, 64, 192, 0, False), (1, 64, 256, 0, False),       # expand
    (1, 384, 64, 0, True), (1, 192, 64, 0, True), (1, 256, 64, 0, True),          # project + residual
    (5, 64, 64, 0, True), (3, 64, 64, 0, False),                                  # tail convs + long skip
    (5, 64, 256, 1, False), (3, 64, 256, 1, False),                               # conv + BN + PixelShuffle
    (3, 64, 16, 2, False),                                                        # conv + BN + PixelUnshuffle
    (5, 64, 3, 0, False), (3, 64, 3, 0, False), (7, 128, 48, 0, False),           # thin outputs, generic
])
@pytest.mark.parametrize('shape', [(1, 8, 16), (2, 22, 38)])
def test_conv_tc_bf16(dev, ks, cin, cout, store, res, shape):
    from ofa_b200 import backend as B
    got, ref = _conv_case(dev, ks, cin, cout, max(cin, 64) if ks > 1 else cin + 64, cout if ks > 1 else cout + 64,
                          store, res, shape, torch.bfloat16, B.IMPL_FAST, seed=11)
    assert got.dtype == torch.bfloat16
    assert relerr(got, ref) < 2 ** -7


def test_conv_tc_fp32_nchw_output(dev):
    """last layer of the nets: bf16 NHWC in, fp32 NCHW out (what the caller receives)."""
    from ofa_b200 import functional as OF, backend as B
    import ofa_b200
    ofa_b200.set_impl(B.IMPL_FAST)
    w = rnd(3, 64, 5, 5, seed=1, scale=0.05)
    x = bf16r(rnd(2, 64, 20, 33, seed=2))
    ref = O.sliced_conv(x, bf16r(w), 3)
    cache = OF.PackedWeightCache()
    xd = x.to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    got = OF.conv_bn_act_infer(xd, w.to(dev), 64, 3, 5, None, B.ACT_NONE, B.STORE_PLAIN, None, cache,
                               out_dtype=torch.float32, out_nchw=True)
    assert got.dtype == torch.float32 and got.is_contiguous()
    assert relerr(got, ref) < 1e-4


@pytest.mark.parametrize('ks,cout', [(5, 3), (3, 3), (3, 5), (5, 1)])
@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
def test_conv_out_rows(dev, ks, cout, dtype):
    """a9, thin output (64 -> 3 at 4x resolution): kx folded into the accumulator columns, rolling row
    ring.  Several 128-pixel strips (W = 300), several 32-row blocks with a ragged last one (H = 70),
    batch 2, fp32 NCHW output as the caller receives it."""
    from ofa_b200 import functional as OF, backend as B
    import ofa_b200
    ofa_b200.set_impl(B.IMPL_AUTO)
    w = rnd(cout + 2, 64, ks, ks, seed=41, scale=0.05)
    x = rnd(2, 64, 70, 300, seed=42).to(dtype).float()
    bn = bn_params(cout + 2, 43, dev)
    ref = ref_bn(O.sliced_conv(x, w.to(dtype).float(), cout), bn, cout)

    class _BN:
        weight, bias, running_mean, running_var, eps = bn['gamma'], bn['beta'], bn['mean'], bn['var'], 1e-5
    xd = x.to(dev).to(dtype).contiguous(memory_format=torch.channels_last)
    got = OF.conv_bn_act_infer(xd, w.to(dev), 64, cout, ks, _BN, B.ACT_NONE, B.STORE_PLAIN, None, OF.PackedWeightCache(),
                               out_dtype=torch.float32, out_nchw=True)
    assert got.dtype == torch.float32 and relerr(got, ref) < 1e-4


@pytest.mark.parametrize('ks,cin', [(5, 3), (3, 3), (3, 1)])
@pytest.mark.parametrize('out', [torch.float16, torch.bfloat16, torch.float32])
def test_conv_stem(dev, ks, cin, out):
    """a9, the stem (3 -> 64 on the user's fp32 NCHW image).  16-bit NHWC output: the im2col tcgen05 kernel (operands
    rounded to the output format, fp32 accumulation); fp32 output: the exact CUDA-core kernel."""
    from ofa_b200 import functional as OF, backend as B
    import ofa_b200
    ofa_b200.set_impl(B.IMPL_AUTO)
    w = rnd(64, cin, ks, ks, seed=51, scale=0.2)
    x = rnd(2, cin, 37, 90, seed=52)
    bn = bn_params(64, 53, dev)
    ref = torch.clamp(ref_bn(O.sliced_conv(x, w, 64), bn, 64), 0, 6)

    class _BN:
        weight, bias, running_mean, running_var, eps = bn['gamma'], bn['beta'], bn['mean'], bn['var'], 1e-5
    got = OF.conv_bn_act_infer(x.to(dev), w.to(dev), cin, 64, ks, _BN, B.ACT_RELU6, B.STORE_PLAIN, None, None, out_dtype=out)
    assert got.dtype == out
    assert relerr(got, ref) < {torch.float32: 1e-5, torch.float16: 2 ** -10, torch.bfloat16: 2 ** -7}[out]


@pytest.mark.parametrize('out', [torch.float16, torch.float32])
def test_conv_stem_16ch_pixelunshuffle(dev, out):
    """a9 + a11: X4's first layer (3x3 3 -> 16, BN, PixelUnshuffle) on the stem kernel, ragged tile edges."""
    from ofa_b200 import functional as OF, backend as B
    import ofa_b200
    ofa_b200.set_impl(B.IMPL_AUTO)
    w = rnd(16, 3, 3, 3, seed=61, scale=0.2)
    x = rnd(2, 3, 38, 90, seed=62)
    bn = bn_params(16, 63, dev)
    ref = O.pixel_unshuffle2(ref_bn(O.sliced_conv(x, w, 16), bn, 16))

    class _BN:
        weight, bias, running_mean, running_var, eps = bn['gamma'], bn['beta'], bn['mean'], bn['var'], 1e-5
    got = OF.conv_bn_act_infer(x.to(dev), w.to(dev), 3, 16, 3, _BN, B.ACT_NONE, B.STORE_PIXELUNSHUFFLE2, None, None, out_dtype=out)
    assert got.dtype == out and tuple(got.shape) == (2, 64, 19, 45)
    assert relerr(got, ref) < (1e-5 if out == torch.float32 else 2 ** -10)


# =================================================================================================
# planar tcgen05 MBConv stages (expand / Toeplitz depthwise / project), each through its C-ABI entry
# =================================================================================================
def _bn_struct(bn):
    from ofa_b200 import backend as B
    return B.OfaBn(bn['gamma'].data_ptr(), bn['beta'].data_ptr(), bn['mean'].data_ptr(), bn['var'].data_ptr(), 1e-5)


def _tdt(code):
    from ofa_b200 import backend as B
    return torch.float16 if code == B.OFA_F16 else torch.bfloat16


@pytest.mark.parametrize('ks', [3, 5, 7])
@pytest.mark.parametrize('half', [True, False])
@pytest.mark.parametrize('shape', [(1, 3, 200, 120), (2, 5, 40, 56), (1, 70, 130, 64), (1, 2, 129, 232)])
def test_dw_planar_toeplitz(dev, ks, half, shape):
    """(a1 + a2 + a5 + a6) tensor-core depthwise on channel planes vs the oracle: ragged tile edges in
    both directions (tiles are 128 rows x 112 columns), several planes per CTA, batch > 1."""
    from ofa_b200 import backend as B
    code = B.OFA_F16 if half else B.OFA_BF16
    n, C, H, W = shape
    w7, m75, m53 = dw_weights(11, dev)
    x = (torch.from_numpy(np.random.RandomState(12).rand(n, C, H, W).astype(np.float32)) * 6).to(_tdt(code))
    filt = O.active_filter(w7, {'7to5_matrix': m75, '5to3_matrix': m53}, [3, 5, 7], C, ks)
    bn = bn_params(384, 13, dev)
    ref = torch.clamp(ref_bn(O.dw_conv(x.float(), filt), bn, C), 0, 6)
    xd = x.to(dev).contiguous()
    yd = torch.full_like(xd, 7.0)
    w7d, m75d, m53d = w7.to(dev), m75.to(dev), m53.to(dev)
    bs = _bn_struct(bn)
    B.check(B.lib().ofa_dw_planar_fwd(xd.data_ptr(), yd.data_ptr(), n, C, H, W, w7d.data_ptr(), 7, m75d.data_ptr(),
                                      m53d.data_ptr(), 1, ks, code, ctypes.byref(bs), B.ACT_RELU6,
                                      torch.cuda.current_stream().cuda_stream))
    assert relerr(yd, ref) < (2 ** -9 if half else 2 ** -6.5)


@pytest.mark.parametrize('mid', [192, 256, 384])
@pytest.mark.parametrize('half,trunk_half', [(True, False), (False, False), (True, True)])
@pytest.mark.parametrize('n,hw', [(1, 1000), (3, 2240)])
def test_expand_project_planar(dev, mid, half, trunk_half, n, hw):
    """(a3 + a4 + a5 + a6 + a8) the two 1x1 convs around the planar intermediate vs the oracle: active
    slices of the FULL weights, partial pixel tiles, mid = 192 (a padded 128-row weight tile)."""
    from ofa_b200 import backend as B
    L = B.lib()
    st = torch.cuda.current_stream().cuda_stream
    code, tcode = (B.OFA_F16 if half else B.OFA_BF16), (B.OFA_F16 if trunk_half else B.OFA_BF16)
    w_exp, w_proj = rnd(384, 64, 1, 1, seed=21, scale=0.2), rnd(64, 384, 1, 1, seed=22, scale=0.1)
    bn1, bn3 = bn_params(384, 23, dev), bn_params(64, 24, dev)
    x = rnd(n, 64, hw, 1, seed=25).to(_tdt(tcode))               # logical NCHW with H = hw, W = 1
    we = torch.empty((mid + 127) // 128 * 128, 64, dtype=_tdt(tcode), device=dev)
    wp = torch.empty(64, mid, dtype=_tdt(code), device=dev)
    wed, wpd = w_exp.to(dev), w_proj.to(dev)
    B.check(L.ofa_mbconv_pack_weights(wed.data_ptr(), wed.stride(0), wed.stride(1), wpd.data_ptr(), wpd.stride(0),
                                      wpd.stride(1), mid, tcode, code, we.data_ptr(), wp.data_ptr(), st))
    # expand: NHWC trunk -> planar
    x_nhwc = x.to(dev).permute(0, 2, 3, 1).contiguous()          # [n, hw, 1, 64]
    t1 = torch.full((n, mid, hw), 7.0, dtype=_tdt(code), device=dev)
    b1 = _bn_struct(bn1)
    B.check(L.ofa_expand_planar_fwd(x_nhwc.data_ptr(), t1.data_ptr(), we.data_ptr(), n, hw, mid, tcode, code,
                                    ctypes.byref(b1), B.ACT_RELU6, st))
    ref1 = torch.clamp(ref_bn(O.sliced_conv(x.float(), w_exp, mid), bn1, mid), 0, 6)
    assert relerr(t1.view(n, mid, hw, 1), ref1) < (2 ** -9 if half and trunk_half else 2 ** -6.5)
    # project (+ residual): planar -> NHWC trunk
    m = (torch.from_numpy(np.random.RandomState(26).rand(n, mid, hw, 1).astype(np.float32)) * 6).to(_tdt(code))
    md = m.to(dev).contiguous()
    y = torch.full((n, hw, 1, 64), 7.0, dtype=_tdt(tcode), device=dev)
    b3 = _bn_struct(bn3)
    for res in (True, False):
        B.check(L.ofa_project_planar_fwd(md.data_ptr(), x_nhwc.data_ptr() if res else None, y.data_ptr(), wp.data_ptr(),
                                         n, hw, mid, tcode, code, ctypes.byref(b3), st))
        ref3 = ref_bn(O.sliced_conv(m.float(), w_proj, 64), bn3, 64) + (x.float() if res else 0)
        assert relerr(y.permute(0, 3, 1, 2), ref3) < (2 ** -9 if half and trunk_half else 2 ** -6.5)


def test_mbconv_planar_shape_sweep_vs_oracle(dev):
    """The planar block (expand -> Toeplitz depthwise -> project + residual) on awkward shapes — single rows and
    columns of tiles, H = 1, one 8-pixel row, batch > 1, sizes straddling the 128-row / 112-column / 256- and
    128-pixel tile edges — against the oracle block (fp32 CPU), fp16 storage."""
    import ofa_b200
    from ofa_b200 import functional as OF, backend as B
    ofa_b200.set_compute_dtype(torch.float16)
    ofa_b200.set_impl(B.IMPL_FAST)       # force the planar path: IMPL_AUTO sends planes this small to the NHWC kernels
    layer = _block_layer(dev).eval()
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1])
    pre = 'blocks.0.mobile_inverted_conv.'
    short = {k[len(pre):]: v for k, v in spec.param_shapes().items() if k.startswith(pre)}
    sd = {pre + k: v for k, v in O.synth_state_dict(short, 21).items()}
    rs = np.random.RandomState(77)
    shapes = [(1, 1, 8), (1, 8, 8), (2, 3, 16), (1, 129, 8), (1, 5, 120), (3, 17, 24), (1, 130, 232), (1, 64, 112), (2, 40, 104)]
    for i, (n, h, w) in enumerate(shapes):
        ks, e = ((7, 6), (5, 4), (3, 3))[i % 3]
        layer.active_kernel_size, layer.active_expand_ratio = ks, e
        x = torch.from_numpy(rs.randn(n, 64, h, w).astype(np.float32)).half().float()
        xd = x.to(dev).half().contiguous(memory_format=torch.channels_last)
        assert OF.planar_supported(xd, 64, 64 * e, 64)
        with torch.no_grad():
            y = layer(xd, xd)                     # block output + identity residual (proxyless_nets.py:50)
            ref = O.mbconv_block(x, sd, 'blocks.0.', ks, e, FULL['ks_list'])
        assert relerr(y, ref) < 4e-3, ((n, h, w), ks, e, relerr(y, ref))


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_mbconv_planar_equals_nhwc_path(dev, dtype):
    """The planar path and the three NHWC kernels are two implementations of the same block
    (ofa_mbconv_fwd picks by shape: small planes go NHWC under IMPL_AUTO): they must
    agree to storage rounding, for every (ks, e); and the AUTO choice follows the documented rule."""
    import ofa_b200
    from ofa_b200 import backend as B
    layer = _block_layer(dev).eval()
    x = rnd(2, 64, 20, 24, seed=33).to(dev).to(dtype).contiguous(memory_format=torch.channels_last)
    for ks in (3, 5, 7):
        for e in (3, 4, 6):
            layer.active_kernel_size, layer.active_expand_ratio = ks, e
            with torch.no_grad():
                ofa_b200.set_impl(B.IMPL_FAST)        # planar even for this small plane (AUTO would pick NHWC)
                y_planar = layer(x)
                ofa_b200.set_impl(B.IMPL_NHWC)
                y_nhwc = layer(x)
            ofa_b200.set_impl(B.IMPL_AUTO)
            assert y_planar.dtype == dtype and relerr(y_planar, y_nhwc) < (2 ** -6 if dtype == torch.bfloat16 else 2 ** -8)
    from ofa_b200 import functional as OF
    mk = lambda h, w: torch.empty(1, 64, h, w, dtype=dtype, device=dev).contiguous(memory_format=torch.channels_last)
    assert not OF.planar_preferred(mk(24, 24)) and OF.planar_preferred(mk(48, 48)) and not OF.planar_preferred(mk(8, 2048))
    assert OF.planar_preferred(mk(96, 96)) and OF.planar_preferred(mk(256, 256)) and OF.planar_preferred(mk(540, 960))


# =================================================================================================
# module level: DynamicMBConvLayer against fixtures of the unmodified reference
# =================================================================================================
def _block_layer(dev):
    from ofa_b200.elastic_nn.modules import DynamicMBConvLayer
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1])
    pre = 'blocks.0.mobile_inverted_conv.'
    short = {k[len(pre):]: v for k, v in spec.param_shapes().items() if k.startswith(pre)}
    layer = DynamicMBConvLayer([64], [64], kernel_size_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], stride=1,
                               act_func='relu6', use_se=False)
    layer.load_state_dict(O.synth_state_dict(short, 21))
    return layer.to(dev)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_block_eval_golden(dev, golden, mode):
    import ofa_b200
    arrays, _ = golden
    layer = _block_layer(dev).eval()
    ofa_b200.set_compute_dtype(torch.float32 if mode == 'fp32' else torch.bfloat16)
    x = torch.from_numpy(arrays['block/x']).to(dev)
    for ks in (3, 5, 7):
        f = layer.depth_conv.conv.get_active_filter(384, ks)
        assert relerr(f, torch.from_numpy(arrays['block/filter_k%d' % ks])) < 1e-5
        for e in (3, 4, 6):
            layer.active_kernel_size, layer.active_expand_ratio = ks, e
            with torch.no_grad():
                xin = x if mode == 'fp32' else x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
                y = layer(xin)
            ref = torch.from_numpy(arrays['block/eval_k%d_e%d' % (ks, e)])
            assert relerr(y, ref) < (1e-4 if mode == 'fp32' else 2e-2)
            # secondary oracle: the static twin built by get_active_subnet computes the same thing
            sub = layer.get_active_subnet(64).eval()
            with torch.no_grad():
                ys = sub(xin)
            assert relerr(ys, y) < (1e-5 if mode == 'fp32' else 2e-2)


@pytest.mark.parametrize('ks,e', [(3, 4), (5, 6), (7, 3)])
def test_block_training_golden(dev, golden, ks, e):
    """a5 (training BN on the slice, running-stat update) + a14 (every gradient of the block)."""
    arrays, _ = golden
    tag = 'block/train_k%d_e%d/' % (ks, e)
    layer = _block_layer(dev).train()
    layer.active_kernel_size, layer.active_expand_ratio = ks, e
    x = torch.from_numpy(arrays['block/x']).to(dev).requires_grad_(True)
    y = layer(x)
    tgt = torch.from_numpy(arrays[tag + 'target']).to(dev)
    loss = torch.nn.functional.mse_loss(y, tgt)
    loss.backward()
    assert relerr(y, torch.from_numpy(arrays[tag + 'y'])) < 1e-4
    assert abs(loss.item() - float(arrays[tag + 'loss'])) < 1e-4 * abs(float(arrays[tag + 'loss']))
    assert relerr(x.grad, torch.from_numpy(arrays[tag + 'dx'])) < 1e-3
    for pname, p in layer.named_parameters():
        ref = arrays[tag + 'grad/' + pname]
        if ref.size == 0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, pname
        else:
            assert p.grad is not None, pname
            assert relerr(p.grad, torch.from_numpy(ref)) < 1e-3, pname
    for bname, b in layer.named_buffers():
        ref = torch.from_numpy(arrays[tag + 'buf/' + bname])
        if b.dtype == torch.int64:
            assert int(b) == int(ref), bname
        else:
            assert relerr(b, ref) < 1e-4, bname


@pytest.mark.parametrize('ks,e', [(3, 3), (5, 4), (7, 6)])
def test_block_training_call_is_bit_identical_to_layerwise_path(dev, ks, e, dtype=torch.bfloat16):
    """ofa_mbconv_train_fwd / _bwd (one library call each way per MBConv block) against the layer-by-layer autograd
    nodes: same kernels in the same order, so outputs, every gradient and the BatchNorm buffers must be EQUAL, over two
    steps (running statistics + num_batches_tracked), with and without the identity shortcut."""
    import ofa_b200
    from ofa_b200 import functional as OF, backend as B
    from ofa_b200.layers import MobileInvertedResidualBlock, IdentityLayer
    ofa_b200.set_train_dtype(dtype)
    try:
        rs = np.random.RandomState(11 * ks + e)
        xs = [torch.from_numpy(rs.randn(6, 64, 24, 24).astype(np.float32)).to(dev).to(dtype)
              .contiguous(memory_format=torch.channels_last) for _ in range(2)]
        dys = [torch.from_numpy(rs.randn(6, 64, 24, 24).astype(np.float32)).to(dev).to(dtype)
               .contiguous(memory_format=torch.channels_last) for _ in range(2)]
        for shortcut in (True, False):
            results = []
            for block_mode in (True, False):
                OF.set_block_train(block_mode)
                layer = _block_layer(dev).train()
                layer.active_kernel_size, layer.active_expand_ratio = ks, e
                mod = MobileInvertedResidualBlock(layer, IdentityLayer(64, 64)) if shortcut else layer
                B.launch_count_reset()
                rec = []
                for x0, dy in zip(xs, dys):
                    x = x0.clone().requires_grad_(True)
                    for p in layer.parameters():
                        p.grad = None
                    y = mod(x)
                    y.backward(dy)
                    rec.append((y.detach().clone(), x.grad.clone(),
                                {n: (None if p.grad is None else p.grad.clone()) for n, p in layer.named_parameters()}))
                results.append((rec, {n: b.clone() for n, b in layer.named_buffers()}, B.launch_count()))
            (ra, ba, la), (rb, bb, lb) = results
            assert la > 0 and lb > 0
            for (ya, dxa, ga), (yb, dxb, gb) in zip(ra, rb):
                assert torch.equal(ya, yb)
                if shortcut:
                    # the block call adds the identity branch's gradient in the fp32 epilogue of the data-gradient conv
                    # (one rounding); the layer-by-layer path rounds the conv result and then adds dy (two roundings)
                    assert relerr(dxa, dxb) < (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10)
                else:
                    assert torch.equal(dxa, dxb)
                assert torch.isfinite(ya.float()).all() and float(dxa.float().abs().max()) > 0
                for n in ga:
                    assert (ga[n] is None) == (gb[n] is None), n
                    if ga[n] is not None:
                        # weight / filter / transform-matrix gradients are summed over pixel splits with fp32 atomics
                        # (the order varies from launch to launch); the BatchNorm reductions are fixed-order
                        if 'bn.' in n:
                            assert torch.equal(ga[n], gb[n]), n
                        else:
                            assert relerr(ga[n], gb[n]) < 1e-5, n
            for n in ba:
                assert torch.equal(ba[n], bb[n]), n
    finally:
        OF.set_block_train(True)
        ofa_b200.set_train_dtype(torch.float32)


def test_block_backward_side_stream_changes_nothing(dev):
    """ofa_train_side_mode(1) only moves the three weight-gradient computations of a block onto a second stream (fork /
    join inside the call): outputs, data gradients and BatchNorm gradients are bit-equal to mode 0, the atomically summed
    weight gradients equal to rounding -- also when the next op on the main stream consumes them immediately."""
    import ofa_b200
    from ofa_b200 import functional as OF
    from ofa_b200.layers import MobileInvertedResidualBlock, IdentityLayer
    ofa_b200.set_train_dtype(torch.bfloat16)
    try:
        rs = np.random.RandomState(77)
        x0 = torch.from_numpy(rs.randn(8, 64, 24, 24).astype(np.float32)).to(dev).to(torch.bfloat16) \
            .contiguous(memory_format=torch.channels_last)
        dy = torch.from_numpy(rs.randn(8, 64, 24, 24).astype(np.float32)).to(dev).to(torch.bfloat16) \
            .contiguous(memory_format=torch.channels_last)
        out = []
        for side in (False, True):
            ofa_b200.set_train_side_stream(side)
            layer = _block_layer(dev).train()
            layer.active_kernel_size, layer.active_expand_ratio = 5, 4
            mod = MobileInvertedResidualBlock(layer, IdentityLayer(64, 64))
            x = x0.clone().requires_grad_(True)
            y = mod(x)
            y.backward(dy)
            # consumed on the main stream right away (no synchronisation in between)
            gsum = sum(p.grad.double().sum() for p in layer.parameters() if p.grad is not None)
            out.append((y.detach().clone(), x.grad.clone(), {n: p.grad.clone() for n, p in layer.named_parameters() if p.grad is not None},
                        float(gsum)))
        (ya, dxa, ga, sa), (yb, dxb, gb, sb) = out
        assert torch.equal(ya, yb) and torch.equal(dxa, dxb)
        assert set(ga) == set(gb)
        for n in ga:
            if 'bn.' in n:
                assert torch.equal(ga[n], gb[n]), n
            else:
                assert relerr(ga[n], gb[n]) < 1e-5, n
        assert abs(sa - sb) <= 1e-4 * max(1.0, abs(sa))
    finally:
        ofa_b200.set_train_side_stream(True)
        ofa_b200.set_train_dtype(torch.float32)


def test_block_training_call_in_network_step(dev):
    """The S4 training step (sampled sub-network, bf16, FusedAdam) with the block-level calls equals the layer-by-layer
    path: loss and every parameter after two steps."""
    import ofa_b200
    from ofa_b200 import functional as OF, optim
    ofa_b200.set_train_dtype(torch.bfloat16)
    try:
        rs = np.random.RandomState(5)
        x = torch.from_numpy(rs.rand(4, 3, 24, 24).astype(np.float32)).to(dev)
        tgt = torch.from_numpy(rs.rand(4, 3, 96, 96).astype(np.float32)).to(dev)
        out = []
        for block_mode in (True, False):
            OF.set_block_train(block_mode)
            net = _build_net('s4', [1, 2], 33, dev).train()
            decay, no_decay = optim.split_no_decay(net.named_parameters())
            opt = optim.FusedAdam(decay, no_decay, lr=1e-3, weight_decay=3e-5)
            losses = []
            for step in range(2):
                random.seed(100 + step)
                net.sample_active_subnet()
                net.set_active_subnet(pixel_d=2)
                net.zero_grad(set_to_none=True)
                loss = torch.nn.functional.mse_loss(net(x), tgt)
                loss.backward()
                opt.step()
                losses.append(float(loss))
            out.append((losses, {n: p.detach().clone() for n, p in net.named_parameters()},
                        {n: b.clone() for n, b in net.named_buffers()}))
        (la, pa, ba), (lb, pb, bb) = out
        # Not bit-equal by construction: the block call rounds the trunk gradient once (identity branch added in the
        # fp32 epilogue of the data-gradient conv), the layer-wise path twice, and weight gradients are summed with
        # fp32 atomics.  The first loss (same weights) is equal; after two Adam steps the loss agrees to 1e-3 and the
        # accumulated parameter updates point the same way.
        assert la[0] == lb[0] and abs(la[1] - lb[1]) <= 1e-3 * abs(lb[1])
        p0 = {n: p.detach().clone() for n, p in _build_net('s4', [1, 2], 33, dev).named_parameters()}
        ua = torch.cat([(pa[n] - p0[n]).flatten() for n in pa]).double()
        ub = torch.cat([(pb[n] - p0[n]).flatten() for n in pa]).double()
        assert float(ua.norm()) > 0
        cos = float(ua @ ub / (ua.norm() * ub.norm()))
        assert cos > 0.98, cos
        assert ((ua == 0) == (ub == 0)).all()          # the same parameters were touched (active-slice bookkeeping)
        for n in ba:
            if ba[n].dtype == torch.int64:
                assert torch.equal(ba[n], bb[n]), n
            else:
                assert relerr(ba[n], bb[n]) < 1e-2, n
    finally:
        OF.set_block_train(True)
        ofa_b200.set_train_dtype(torch.float32)


# =================================================================================================
# network level: S4 / X4 against the reference fixtures — fp32 exact path and bf16 fast path
# =================================================================================================
def _build_net(kind, pd, wseed, dev):
    from ofa_b200.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
    cls = OFAMobileNetS4 if kind == 's4' else OFAMobileNetX4
    net = cls(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=list(pd))
    spec = O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], pd)
    net.load_state_dict(O.synth_state_dict(spec.param_shapes(), wseed))
    return net.to(dev).eval()


def _apply(net, request):
    if isinstance(request, str):
        random.seed(int(request.split(':')[1]))
        return net.sample_active_subnet()
    net.set_active_subnet(**request)
    return dict(request)


@pytest.mark.parametrize('name', ['s4_ps1', 's4_ps12', 'x4_ps12'])
@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_net_forward_golden(dev, golden, name, mode):
    import ofa_b200
    arrays, book = golden
    meta = book[name]
    net = _build_net(meta['kind'], meta['pd'], meta['wseed'], dev)
    ofa_b200.set_compute_dtype(torch.float32 if mode == 'fp32' else torch.bfloat16)
    x = torch.from_numpy(arrays[name + '/x']).to(dev)
    for i, sub in enumerate(meta['subnets']):
        setting = _apply(net, sub['request'])
        assert setting == sub['setting'] and list(net.runtime_depth) == sub['runtime_depth']
        with torch.no_grad():
            y = net(x)
        ref = torch.from_numpy(arrays['%s/y%d' % (name, i)])
        assert list(y.shape) == sub['out_shape'] and y.dtype == torch.float32
        if mode == 'fp32':
            assert relerr(y, ref) < 1e-3
        else:
            assert relerr(y, ref) < 5e-2      # the PSNR criterion needs image-sized outputs: next test


@pytest.mark.parametrize('kind,shape', [('s4', (1, 3, 96, 96)), ('s4', (1, 3, 64, 64)), ('x4', (1, 3, 256, 256))])
@pytest.mark.parametrize('storage', ['fp16', 'bf16'])
def test_16bit_psnr_within_0p01_db(dev, kind, shape, storage):
    """north_star: 'PSNR within 0.01 dB' for the 16-bit tensor-core path.  PSNR is a statistic over
    pixels, so it is evaluated on image-sized outputs (256x256, >= 65k pixels).  Weights are the plain
    O(1) synthetic recipe (no down-scaled branches).  The 96x96 LR input runs the planar frame path (planes >= 8192
    pixels), the smaller ones the NHWC kernels.  fp16 storage (the default) is held to the 0.01 dB
    of the north star; bf16 storage to 0.03 dB: rounding the conv OPERANDS to bf16 already costs
    0.012-0.019 dB on these nets even with fp32 storage everywhere (CPU emulation), so
    0.01 dB is not reachable by any bf16 tensor-core pipeline at this depth."""
    import ofa_b200
    ofa_b200.set_compute_dtype(torch.float16 if storage == 'fp16' else torch.bfloat16)
    net = _build_net(kind, [1, 2], 61, dev)
    spec = O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
    sd = O.synth_state_dict(spec.param_shapes(), 61)
    net.load_state_dict(sd)
    x = torch.from_numpy(np.random.RandomState(8).rand(*shape).astype(np.float32))
    bound = 0.01 if storage == 'fp16' else 0.03
    for sub in (dict(ks=7, e=6, d=4, pixel_d=2), dict(ks=3, e=3, d=2, pixel_d=1), dict(ks=5, e=4, d=3, pixel_d=2)):
        net.set_active_subnet(**sub)
        spec.set_active_subnet(**sub)
        with torch.no_grad():
            y = net(x.to(dev)).cpu()
            ref = O.supernet_forward(x, sd, spec)
        assert y.shape == ref.shape
        d = psnr_delta_db(ref, y)
        print('psnr delta %s %s %s: %.4f dB' % (storage, kind, sub, d))
        assert d < bound, (storage, kind, sub, d)


def psnr_delta_db(ref, got, target_psnr_db=31.0):
    """|PSNR(got, T) - PSNR(ref, T)| with the reference's metric (clamp -> uint8 -> BT.601 Y -> PSNR,
    sr_run_manager.py:567-597) in an SR-like operating point: random-weight outputs have no meaningful
    range, so both images go through the SAME affine map that puts the reference's 1..99 percentile on
    [0.1, 0.9], and the target T is the mapped reference plus Gaussian noise at `target_psnr_db`
    (31 dB = the reference README's 4x Set14 figure)."""
    lo, hi = np.percentile(ref.numpy(), [1, 99])
    a = 0.8 / max(hi - lo, 1e-6)
    f = lambda t: (t - lo) * a + 0.1
    sigma = 1.0 / (10 ** (target_psnr_db / 20.0))
    worst = 0.0
    for b in range(ref.shape[0]):
        r, g = f(ref[b]), f(got[b])
        noise = torch.from_numpy(np.random.RandomState(17 + b).randn(*r.shape).astype(np.float32)) * sigma
        t = O.tensor_to_y_uint8(r + noise)
        worst = max(worst, abs(O.psnr_uint8(O.tensor_to_y_uint8(r), t) - O.psnr_uint8(O.tensor_to_y_uint8(g), t)))
    return worst


def test_random_subnet_sweep_vs_oracle(dev):
    """C5: 200 sampled subnets (100 per net, seeded through Python `random` exactly as
    progressive_shrinking.py:164 does) — selection bit-exact, outputs vs the oracle on the exact fp32 path;
    every 10th subnet also through the fp16 tensor-core path."""
    import ofa_b200
    for kind, pd, shape in (('s4', [1, 2], (1, 3, 10, 16)), ('x4', [1, 2], (1, 3, 16, 32))):
        net = _build_net(kind, pd, 41, dev)
        spec = O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], pd)
        sd = O.synth_state_dict(spec.param_shapes(), 41)
        x = torch.from_numpy(np.random.RandomState(3).rand(*shape).astype(np.float32))
        for seed in range(100):
            random.seed(seed)
            a = net.sample_active_subnet()
            random.seed(seed)
            b = spec.sample_active_subnet()
            assert a == b and list(net.runtime_depth) == spec.runtime_depth
            with torch.no_grad():
                ref = O.supernet_forward(x, sd, spec)
                ofa_b200.set_compute_dtype(torch.float32)
                y = net(x.to(dev))
                assert relerr(y, ref) < 1e-3, (kind, seed)
                if seed % 10 == 0:
                    ofa_b200.set_compute_dtype(torch.float16)
                    y16 = net(x.to(dev))
                    assert relerr(y16, ref) < 1e-2, (kind, seed)


@pytest.mark.parametrize('kind', ['s4', 'x4'])
def test_mixed_precision_training_step(dev, kind):
    """C3 as BASELINE.json names it (bf16 compute, fp32 master weights): forward convs and data gradients on
    the tcgen05 kernel, bf16 activations / activation gradients, fp32 weight gradients.  Against the oracle's
    fp32 CPU autograd: loss within 2 %, conv-weight gradient norms within 6 %, directions (cosine) > 0.95 (S4) /
    0.90 (X4) per conv weight and > 0.999 next to the loss, whole-gradient cosine > 0.93 and norm within 5 %."""
    import ofa_b200
    ofa_b200.set_train_dtype(torch.bfloat16)
    try:
        net = _build_net(kind, [1, 2], 81, dev).train()
        spec = O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
        sd = O.synth_state_dict(spec.param_shapes(), 81)
        sd_ref = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and 'running' not in k else v.clone())
                  for k, v in sd.items()}
        rs = np.random.RandomState(6)
        if kind == 's4':
            x = torch.from_numpy(rs.rand(4, 3, 12, 16).astype(np.float32))
            tgt = torch.from_numpy(rs.rand(4, 3, 48, 64).astype(np.float32))
        else:
            x = torch.from_numpy(rs.rand(4, 3, 32, 48).astype(np.float32))
            tgt = x.clone()
        sub = dict(ks=7, e=6, d=4, pixel_d=2) if kind == 's4' else dict(ks=5, e=4, d=3, pixel_d=2)
        net.set_active_subnet(**sub)
        spec.set_active_subnet(**sub)
        net.zero_grad()
        out = net(x.to(dev))
        assert out.dtype == torch.float32
        loss = torch.nn.functional.mse_loss(out, tgt.to(dev))
        loss.backward()
        out_ref = O.supernet_forward(x, sd_ref, spec, training=True)
        loss_ref = torch.nn.functional.mse_loss(out_ref, tgt)
        loss_ref.backward()
        assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 2e-2 * float(loss_ref.detach())
        checked, dot, n_a, n_b = 0, 0.0, 0.0, 0.0
        for pname, p in net.named_parameters():
            g_ref = sd_ref[pname].grad
            if g_ref is None or float(g_ref.norm()) == 0.0:
                assert p.grad is None or float(p.grad.norm()) <= 1e-6, pname     # participation is exact
                continue
            assert p.grad is not None and p.grad.dtype == torch.float32, pname
            ga, gb = p.grad.cpu().flatten().double(), g_ref.flatten().double()
            dot, n_a, n_b = dot + float(ga @ gb), n_a + float(ga @ ga), n_b + float(gb @ gb)
            if g_ref.numel() >= 1024:
                # conv weights.  The deviation from the fp32 oracle grows smoothly with depth (tools/dbg_train_x4.py:
                # 0.9996 at the last conv down to ~0.92 behind X4's 60+ layers); it is not rounding of single ops
                # (each is at 0.999996 against its fp32 twin, tools/dbg_train.py) but ReLU6 masks and batch
                # statistics being evaluated on bf16-rounded activations, as in any bf16 mixed-precision training.
                n_ref = float(g_ref.norm())
                assert abs(float(p.grad.norm()) - n_ref) <= 6e-2 * n_ref, (pname, float(p.grad.norm()), n_ref)
                cos = float(ga @ gb) / (float(ga.norm()) * float(gb.norm()))
                assert cos > (0.95 if kind == 's4' else 0.90), (pname, cos)
                if pname.startswith('dec_final_output_conv_block'):
                    assert cos > 0.999, (pname, cos)
            checked += 1
        # the whole gradient (what the optimizer sees): direction and length
        assert dot / (n_a ** 0.5 * n_b ** 0.5) > 0.93
        assert abs(n_a ** 0.5 - n_b ** 0.5) <= 5e-2 * n_b ** 0.5
        assert checked > 60
    finally:
        ofa_b200.set_train_dtype(torch.float32)


def test_x4_joint_distillation_step(dev):
    """C4: the task-aware downscale -> upscale net trained at 2x and 4x in the same step with teacher
    distillation (progressive_shrinking.py:158-203, kd_type != 'ce' branch): teacher = the max subnet under
    no_grad, loss = MSE(out, HR) + kd_ratio * MSE(out, teacher_out), gradients accumulated over the sampled
    subnets.  Product (CUDA autograd Functions) vs the oracle (torch CPU autograd on the same weights)."""
    import ofa_b200
    ofa_b200.set_compute_dtype(torch.float32)
    net = _build_net('x4', [1, 2], 71, dev).train()
    spec = O.SuperNetSpec('x4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
    sd = O.synth_state_dict(spec.param_shapes(), 71)
    sd_ref = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and 'running' not in k else v.clone())
              for k, v in sd.items()}
    hr = torch.from_numpy(np.random.RandomState(5).rand(2, 3, 16, 16).astype(np.float32))
    hr_d = hr.to(dev)
    kd = 0.5
    net.zero_grad()
    # teacher outputs (max subnet, eval-mode BN like the reference's frozen teacher copy), both scales
    teach, teach_ref = {}, {}
    net.eval()
    for pdepth in (1, 2):
        net.set_active_subnet(ks=7, e=6, d=4, pixel_d=pdepth)
        spec.set_active_subnet(ks=7, e=6, d=4, pixel_d=pdepth)
        with torch.no_grad():
            teach[pdepth] = net(hr_d)
            teach_ref[pdepth] = O.supernet_forward(hr, {k: v.detach() for k, v in sd_ref.items()}, spec)
        assert relerr(teach[pdepth], teach_ref[pdepth]) < 1e-3
    net.train()
    losses, losses_ref = [], []
    for sub in (dict(ks=5, e=4, d=3, pixel_d=1), dict(ks=3, e=6, d=2, pixel_d=2)):   # a 2x and a 4x student
        pdepth = sub['pixel_d']
        net.set_active_subnet(**sub)
        spec.set_active_subnet(**sub)
        assert list(net.runtime_depth) == spec.runtime_depth
        out = net(hr_d)
        loss = torch.nn.functional.mse_loss(out, hr_d) + kd * torch.nn.functional.mse_loss(out, teach[pdepth])
        loss.backward()
        out_ref = O.supernet_forward(hr, sd_ref, spec, training=True)
        loss_ref = torch.nn.functional.mse_loss(out_ref, hr) + kd * torch.nn.functional.mse_loss(out_ref, teach_ref[pdepth])
        loss_ref.backward()
        losses.append(float(loss.detach()))
        losses_ref.append(float(loss_ref.detach()))
    np.testing.assert_allclose(losses, losses_ref, rtol=2e-3)
    checked = 0
    for pname, p in net.named_parameters():
        g_ref = sd_ref[pname].grad
        if g_ref is None or float(g_ref.norm()) == 0.0:
            assert p.grad is None or float(p.grad.norm()) <= 1e-6, pname
            continue
        assert p.grad is not None, pname
        n_ref = float(g_ref.norm())
        assert abs(float(p.grad.norm()) - n_ref) <= 5e-3 * n_ref, (pname, float(p.grad.norm()), n_ref)
        checked += 1
    assert checked > 100


def test_s4_training_step_golden(dev, golden):
    """C3 shape of work: two sampled subnets, gradients accumulate, BN in batch-stat mode."""
    arrays, book = golden
    net = _build_net('s4', [1, 2], 31, dev).train()
    lr_img = torch.from_numpy(arrays['train_s4/lr']).to(dev)
    hr_img = torch.from_numpy(arrays['train_s4/hr']).to(dev)
    net.zero_grad()
    losses = []
    for j in range(2):
        random.seed(int('%d%.3d%.3d' % (5, j, 0)))
        net.sample_active_subnet()
        y = net(lr_img)
        loss = torch.nn.functional.mse_loss(y, hr_img)
        loss.backward()
        losses.append(loss.item())
    assert list(net.runtime_depth) == book['train_s4_runtime_depth']
    np.testing.assert_allclose(losses, arrays['train_s4/losses'], rtol=1e-3)
    for pname, p in net.named_parameters():
        ref = book['train_s4_grad_norms'][pname]
        if ref is None:
            assert p.grad is None or float(p.grad.norm()) == 0.0, pname   # inactive block / matrix
        else:
            assert p.grad is not None, pname
            assert abs(float(p.grad.norm()) - ref) <= 2e-3 * max(ref, 1e-6), (pname, float(p.grad.norm()), ref)
    for key in ('dec_first_conv_block.conv.weight', 'blocks.0.mobile_inverted_conv.depth_conv.conv.conv.weight'):
        g = dict(net.named_parameters())[key].grad
        assert relerr(g, torch.from_numpy(arrays['train_s4/grad/' + key])) < 2e-3, key
    rv = net.blocks[0].mobile_inverted_conv.depth_conv.bn.bn.running_var
    assert relerr(rv, torch.from_numpy(arrays['train_s4/buf/blocks.0.mobile_inverted_conv.depth_conv.bn.bn.running_var'])) < 1e-4


# =================================================================================================
# edge cases and error behaviour
# =================================================================================================
def test_ragged_and_tiny_images(dev):
    import ofa_b200
    net = _build_net('s4', [1, 2], 51, dev)
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
    sd = O.synth_state_dict(spec.param_shapes(), 51)
    net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
    spec.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
    for shape in ((1, 3, 1, 1), (1, 3, 5, 3), (3, 3, 9, 17)):
        x = torch.from_numpy(np.random.RandomState(1).rand(*shape).astype(np.float32))
        ref = O.supernet_forward(x, sd, spec)
        for dt, tol in ((torch.float32, 1e-3), (torch.bfloat16, 5e-2)):
            ofa_b200.set_compute_dtype(dt)
            with torch.no_grad():
                y = net(x.to(dev))
            assert relerr(y, ref) < tol, (shape, dt)


def test_device_psnr_metric_is_bit_exact(dev):
    """§8f-1: the validate metric on the device — integer SSE bit-exact against the oracle (which is pinned to the
    reference's own functions), PSNR equal to the reference's golden values incl. the make_grid batch quirk;
    fp32 NCHW, channels-last and 16-bit inputs."""
    import json
    import math
    import os
    from ofa_b200 import metrics
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_metric.json')) as f:
        cases = json.load(f)
    for c in cases:
        rs = np.random.RandomState(c['seed'])
        a = (rs.rand(*c['shape']) * 1.2 - 0.1).astype(np.float32)
        b = (a + c['noise'] * rs.randn(*c['shape'])).astype(np.float32)
        ta, tb = torch.from_numpy(a), torch.from_numpy(b)
        sse = metrics.psnr_y_sse(ta.to(dev), tb.to(dev).contiguous(memory_format=torch.channels_last))
        assert sse.cpu().tolist() == O.psnr_y_sse(ta, tb).tolist()
        got = metrics.psnr_y(ta.to(dev), tb.to(dev))
        if c['psnr'] is None:
            assert math.isinf(got)
        else:
            assert abs(got - c['psnr']) < 1e-9
        h16a, h16b = ta.half(), tb.half()
        assert metrics.psnr_y_sse(h16a.to(dev), h16b.to(dev)).cpu().tolist() == O.psnr_y_sse(h16a.float(), h16b.float()).tolist()


def test_bn_recalibration_vs_oracle(dev):
    """§8f-2: elastic_nn.utils.set_running_statistics on the product (statistics and normalisation on the device, the
    fused inference kernels step aside for the per-BatchNorm forward overrides) vs the oracle, which is pinned to the
    reference's own function.  Inactive blocks / channels must keep their old statistics bit for bit."""
    import ofa_b200
    from ofa_b200.elastic_nn.utils import set_running_statistics
    ofa_b200.set_compute_dtype(torch.float32)
    net = _build_net('s4', [1, 2], 95, dev)
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
    sd = O.synth_state_dict(spec.param_shapes(), 95)
    before = {k: v.clone() for k, v in sd.items()}
    sub = dict(ks=5, e=4, d=3, pixel_d=2)
    net.set_active_subnet(**sub)
    spec.set_active_subnet(**sub)
    rs = np.random.RandomState(9)
    batches = [torch.from_numpy(rs.rand(3, 3, 8, 12).astype(np.float32)), torch.from_numpy(rs.rand(2, 3, 8, 12).astype(np.float32))]
    set_running_statistics(net, [{'image': b} for b in batches])
    O.set_running_statistics(sd, spec, batches)
    got = net.state_dict()
    changed = 0
    for k, v in sd.items():
        if k.endswith('running_mean') or k.endswith('running_var'):
            np.testing.assert_allclose(got[k].cpu().numpy(), v.numpy(), rtol=2e-4, atol=2e-6, err_msg=k)
            same = torch.equal(v, before[k])
            assert same == torch.equal(got[k].cpu(), before[k]), k       # untouched buffers stay untouched
            changed += 0 if same else 1
    assert changed > 40
    # and the re-calibrated net still runs through the fused inference path
    with torch.no_grad():
        y = net(batches[0].to(dev))
    assert relerr(y, O.supernet_forward(batches[0], sd, spec)) < 1e-3


def test_fused_adam_matches_torch_adam(dev):
    """§8f-3: FusedAdam vs torch.optim.Adam built as sr_run_manager.py:115-133 builds it (two groups, L2 weight decay
    except on the 'bn#bias' keys): five steps with a cosine learning rate, parameters that have no gradient in some
    steps (inactive blocks) must be skipped — moments, step counters and weights untouched."""
    from ofa_b200 import optim
    torch.manual_seed(3)
    shapes = {'conv.weight': (64, 3, 5, 5), 'bn.weight': (64,), 'bn.bias': (64,), 'blocks.1.conv.weight': (384, 64, 1, 1),
              'blocks.1.bn.weight': (384,), 'blocks.2.conv.weight': (5000,), 'blocks.2.bias': (7,)}
    ref_p = {k: torch.randn(*v).requires_grad_(True) for k, v in shapes.items()}
    our_p = {k: v.detach().clone().to(dev).requires_grad_(True) for k, v in ref_p.items()}
    d_ref, nd_ref = optim.split_no_decay(ref_p.items())
    d_our, nd_our = optim.split_no_decay(our_p.items())
    assert len(nd_our) == 4
    ref = torch.optim.Adam([{'params': d_ref, 'weight_decay': 3e-5}, {'params': nd_ref, 'weight_decay': 0}], 1e-2)
    ours = optim.FusedAdam(d_our, nd_our, lr=1e-2, weight_decay=3e-5)
    for step in range(5):
        lr = optim.cosine_lr(1e-2, 2, step // 3, step % 3, 3)
        for g in ref.param_groups:
            g['lr'] = lr
        ours.set_lr(lr)
        for k in shapes:
            active = not (k.startswith('blocks.2') and step in (1, 3))      # an inactive block in two of the steps
            g = torch.randn(*shapes[k]) if active else None
            ref_p[k].grad = g
            our_p[k].grad = None if g is None else g.to(dev)
        ref.step()
        ours.step()
        for k in shapes:
            assert relerr(our_p[k], ref_p[k]) < 2e-6, (step, k)
    assert ours._steps.cpu().tolist().count(3) == 2 and ours._steps.cpu().tolist().count(5) == 5
    assert abs(optim.warmup_lr(0.1, 50, 10, 1, 3, 0.01) - ((1 * 10 + 3 + 1) / 50 * (0.1 - 0.01) + 0.01)) < 1e-12


def test_tiled_inference_equals_whole_frame(dev):
    """§8e: spatial tiles with a 64-pixel LR halo (>= the max sub-network's 51.5-pixel receptive-field radius) are
    independent units — the tile grid of ofa_b200.parallel reproduces the whole-frame result (eval-mode BN is a
    per-channel affine), here for all 4 'ranks' of a 2 x 2 grid on one GPU; no collective is involved."""
    import ofa_b200
    from ofa_b200 import parallel as P
    ofa_b200.set_compute_dtype(torch.float16)
    net = _build_net('s4', [1, 2], 53, dev)
    net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
    x = torch.rand(1, 3, 160, 176, device=dev)
    with torch.no_grad():
        whole = net(x)
        out = None
        for rank in range(4):
            out, mine = P.tiled_forward(net, x, 2, 2, halo=P.S4_HALO_LR, scale=4, rank=rank, world=4, out=out)
            assert len(mine) == 1
    # identical arithmetic per pixel up to the order in which the depthwise tiles accumulate their column chunks
    assert relerr(out, whole) < 2e-3


def test_empty_batch_returns_empty(dev):
    """N = 0 (the reference's F.conv2d path returns an empty tensor of the right shape)."""
    import ofa_b200
    net = _build_net('s4', [1, 2], 51, dev)
    net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
    for dt in (torch.float16, torch.float32):
        ofa_b200.set_compute_dtype(dt)
        with torch.no_grad():
            y = net(torch.rand(0, 3, 16, 24, device=dev))
        assert tuple(y.shape) == (0, 3, 64, 96) and y.dtype == torch.float32


def test_cuda_graph_capture_is_bit_identical(dev):
    """The whole inference forward can be captured in a CUDA graph (every kernel launches on the caller's
    current stream, tensor maps are kernel parameters) and replays bit-identically."""
    import ofa_b200
    ofa_b200.set_compute_dtype(torch.float16)
    net = _build_net('s4', [1, 2], 52, dev)
    net.set_active_subnet(ks=5, e=4, d=3, pixel_d=2)
    x = torch.rand(1, 3, 96, 104, device=dev)        # planes >= 8192 pixels: the planar frame path is what gets captured
    with torch.no_grad():
        y = net(x)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net(x)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            yg = net(x)
        x.copy_(torch.rand_like(x))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(yg, net(x)) and not torch.equal(yg, y)


def test_no_cpu_path(dev):
    from ofa_b200 import functional as OF
    w7, m75, m53 = dw_weights(1, dev)
    with pytest.raises(RuntimeError):
        OF.dw_conv(torch.randn(1, 64, 4, 4), w7, m75, m53, 3, True)


def test_bad_arguments_are_reported(dev):
    from ofa_b200 import backend as B
    x = torch.randn(1, 64, 4, 4, device=dev)
    tx = B.t4(x)
    w = torch.randn(64, 49, device=dev)
    rc = B.lib().ofa_dw_fwd(ctypes.byref(tx), ctypes.byref(tx), w.data_ptr(), 7, None, None, 1, 4, None, 0, None)
    assert rc == 1 and b'kernel size' in B.lib().ofa_last_error()
    with pytest.raises(RuntimeError):
        B.check(rc)


# =================================================================================================
# BASELINE.json's full sizes, through size-independent properties
# =================================================================================================
def test_full_size_frame_c2_window_vs_oracle_and_tiling(dev):
    """configs[1] at its full size: S4 max sub-network, LR 960x540 -> 3840x2160, fp16 storage.
    (1) Locality: in eval mode BatchNorm is a per-channel affine, so the SR output over a window depends only on the
        LR pixels within the receptive field (< 64 LR pixels).  The oracle runs on a 224x224 LR crop (96x96 window +
        64 halo) -- seconds on the CPU -- and must agree with that window of the full-frame CUDA result; one window
        in the interior, one in the bottom-right corner (ragged depthwise tiles, zero padding).
    (2) The frame computed as 2 x 4 tiles with halo equals the frame computed whole."""
    import ofa_b200
    from ofa_b200 import parallel as P
    ofa_b200.set_compute_dtype(torch.float16)
    net = _build_net('s4', [1, 2], 61, dev)
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
    sd = O.synth_state_dict(spec.param_shapes(), 61)
    sub = dict(ks=7, e=6, d=4, pixel_d=2)
    net.set_active_subnet(**sub)
    spec.set_active_subnet(**sub)
    H, W, halo, win = 540, 960, P.S4_HALO_LR, 96
    x = torch.from_numpy(np.random.RandomState(12).rand(1, 3, H, W).astype(np.float32))
    with torch.no_grad():
        y = net(x.to(dev))
        assert tuple(y.shape) == (1, 3, 4 * H, 4 * W) and bool(torch.isfinite(y).all())
        for (h0, w0) in ((208, 400), (H - win, W - win)):
            i0, i1 = max(0, h0 - halo), min(H, h0 + win + halo)
            j0, j1 = max(0, w0 - halo), min(W, w0 + win + halo)
            ref = O.supernet_forward(x[:, :, i0:i1, j0:j1], sd, spec)
            ref_win = ref[:, :, 4 * (h0 - i0):4 * (h0 - i0 + win), 4 * (w0 - j0):4 * (w0 - j0 + win)]
            got_win = y[:, :, 4 * h0:4 * (h0 + win), 4 * w0:4 * (w0 + win)]
            assert relerr(got_win, ref_win) < 1e-2, (h0, w0)
        out = None
        for rank in range(8):
            out, mine = P.tiled_forward(net, x.to(dev), 2, 4, halo=halo, scale=4, rank=rank, world=8, out=out)
            assert len(mine) == 1
        assert relerr(out, y) < 2e-3


def test_full_size_training_step_c3_bf16_vs_exact_path(dev):
    """configs[2] at its full size: batch 64 of 96x96 HR patches (24x24 LR in), S4 max sub-network, one
    forward + backward.  The bf16 tensor-core path against the exact fp32 CUDA-core path of the same library (which
    the small-size tests pin to the oracle): loss within 2 %, whole-gradient cosine > 0.93 and norm within 5 %,
    BatchNorm running statistics within 1 %; and the gradient agrees (cosine > 0.99) whether the batch is fed NCHW or
    channels-last."""
    import ofa_b200
    rs = np.random.RandomState(21)
    x = torch.from_numpy(rs.rand(64, 3, 24, 24).astype(np.float32)).to(dev)
    tgt = torch.from_numpy(rs.rand(64, 3, 96, 96).astype(np.float32)).to(dev)
    results = {}
    try:
        for mode, dt in (('fp32', torch.float32), ('bf16', torch.bfloat16), ('bf16_cl', torch.bfloat16)):
            ofa_b200.set_train_dtype(dt)
            net = _build_net('s4', [1, 2], 83, dev).train()
            net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
            xin = x.contiguous(memory_format=torch.channels_last) if mode == 'bf16_cl' else x
            loss = torch.nn.functional.mse_loss(net(xin), tgt)
            loss.backward()
            grads = torch.cat([p.grad.flatten().double() for p in net.parameters() if p.grad is not None])
            stats = torch.cat([b.flatten().double() for n, b in net.named_buffers() if n.endswith('running_var')])
            results[mode] = (float(loss.detach()), grads, stats)
    finally:
        ofa_b200.set_train_dtype(torch.float32)
    l32, g32, s32 = results['fp32']
    l16, g16, s16 = results['bf16']
    assert abs(l16 - l32) <= 2e-2 * l32
    assert g16.shape == g32.shape
    cos = float(g16 @ g32) / (float(g16.norm()) * float(g32.norm()))
    assert cos > 0.93, cos
    assert abs(float(g16.norm()) - float(g32.norm())) <= 5e-2 * float(g32.norm())
    assert float((s16 - s32).abs().max() / s32.abs().max()) < 1e-2
    l_cl, g_cl, _ = results['bf16_cl']
    assert abs(l_cl - l16) <= 1e-3 * l16
    # the two runs differ only by the order of the fp32 atomic accumulations in the weight-gradient kernels, but a bf16
    # rounding that flips behind 14 blocks moves ReLU6 masks: 0.998 - 1.0 observed run to run
    assert float(g_cl @ g16) / (float(g_cl.norm()) * float(g16.norm())) > 0.99


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize('shape', [(2, 256, 5, 7), (1, 64, 12, 24), (3, 8, 3, 2), (1, 12, 4, 6)])
def test_pixel_reorder_kernels_exact(dev, dtype, shape):
    """(a10, a11) stand-alone PixelShuffle(2) / PixelUnshuffle(2) (the training path's reorder, incl. the 16-byte vector
    kernels for 16-bit dense NHWC tensors): pure data movement, bit-exact against torch, and inverse of each other."""
    from ofa_b200 import functional as OF, backend as B
    n, c, h, w = shape
    x = torch.randn(n, c, h, w, device=dev).to(dtype).contiguous(memory_format=torch.channels_last)
    up = OF.ReorderFn.apply(x, B.STORE_PIXELSHUFFLE2)
    assert torch.equal(up, torch.nn.functional.pixel_shuffle(x, 2))
    assert torch.equal(OF.ReorderFn.apply(up, B.STORE_PIXELUNSHUFFLE2), x)
    if h % 2 == 0 and w % 2 == 0:
        down = OF.ReorderFn.apply(x, B.STORE_PIXELUNSHUFFLE2)
        assert torch.equal(down, torch.nn.functional.pixel_unshuffle(x, 2))


def test_frame_width_not_multiple_of_8_is_split_into_column_tiles(dev):
    """A frame whose width is not a multiple of 8 cannot take the planar kernels (16-byte TMA row pitch); in inference
    OFAMobileNetS4 computes it as two column tiles of legal widths that overlap by the receptive field.  The result must
    equal the whole frame computed on the NHWC kernels, and the oracle."""
    import ofa_b200
    from ofa_b200 import backend as B
    ofa_b200.set_compute_dtype(torch.float16)
    net = _build_net('s4', [1, 2], 67, dev)
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
    sd = O.synth_state_dict(spec.param_shapes(), 67)
    x = torch.from_numpy(np.random.RandomState(4).rand(1, 3, 72, 300).astype(np.float32))
    assert net._column_split(x.to(dev)) is None            # grad mode on: the autograd path is never split
    for sub in (dict(ks=7, e=6, d=4, pixel_d=2), dict(ks=3, e=4, d=2, pixel_d=1)):
        net.set_active_subnet(**sub)
        spec.set_active_subnet(**sub)
        with torch.no_grad():
            c, a, b = net._column_split(x.to(dev))
            assert a % 8 == 0 and (300 - b) % 8 == 0 and a - c >= 64 and c - b >= 64
            y = net(x.to(dev))
            ofa_b200.set_impl(B.IMPL_NHWC)                  # whole frame on the NHWC kernels: no split
            assert net._column_split(x.to(dev)) is None
            y_whole = net(x.to(dev))
            ofa_b200.set_impl(B.IMPL_AUTO)
            ref = O.supernet_forward(x, sd, spec)
        assert y.shape == y_whole.shape == ref.shape
        assert relerr(y, y_whole) < 2e-3 and relerr(y, ref) < 1e-2


def test_fused_adam_invalidates_packed_weight_caches(dev):
    """FusedAdam writes the parameters through raw pointers; the derived 16-bit weight copies of the tensor-core convs
    are cached per Tensor._version, so the step must bump the versions — otherwise the next forward (training in bf16,
    or inference) keeps computing with the weights from before the step.  A few large-lr steps on one batch must
    (1) bump every updated parameter's version, (2) change the training forward, (3) reduce the loss, and (4) leave the
    eval-mode fp16 forward consistent with the fp32 exact path on the UPDATED weights."""
    import ofa_b200
    from ofa_b200 import optim
    ofa_b200.set_train_dtype(torch.bfloat16)
    try:
        net = _build_net('s4', [1, 2], 88, dev).train()
        net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
        decay, no_decay = optim.split_no_decay(net.named_parameters())
        opt = optim.FusedAdam(decay, no_decay, lr=2e-3, weight_decay=0.0)
        rs = np.random.RandomState(31)
        x = torch.from_numpy(rs.rand(4, 3, 16, 16).astype(np.float32)).to(dev)
        tgt = torch.from_numpy(rs.rand(4, 3, 64, 64).astype(np.float32)).to(dev)
        w = net.blocks[0].mobile_inverted_conv.inverted_bottleneck.conv.conv.weight
        losses = []
        for step in range(6):
            net.zero_grad(set_to_none=True)
            v0 = w._version
            loss = torch.nn.functional.mse_loss(net(x), tgt)
            loss.backward()
            opt.step()
            assert w._version > v0
            losses.append(float(loss.detach()))
        assert losses[-1] < 0.73 * losses[0], losses         # 0.67 measured; 0.79 with stale forward weights
        net.eval()
        with torch.no_grad():
            ofa_b200.set_compute_dtype(torch.float16)
            y16 = net(x)
            ofa_b200.set_compute_dtype(torch.float32)
            y32 = net(x)
        assert relerr(y16, y32) < 1e-2
    finally:
        ofa_b200.set_train_dtype(torch.float32)
        ofa_b200.set_compute_dtype(torch.bfloat16)
