"""Round-2 GPU tests (`pytest -m gpu`): derived-weight freshness under every writer, the single-launch band-scheduled
block, small-plane batches on the planar path, PixelShuffle widths the coalesced epilogue used to break, sampled
sub-networks on the planar tcgen05 path, 16-bit range / trajectory evidence, optimizer state round trip.

Tolerances: fp32 path <= 1e-3 max relative error; fp16 tensor-core path per net <= 1e-2 of max|ref|; band kernel and
three-kernel planar path: BIT-identical (same MMAs in the same order)."""
import copy
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
    ofa_b200.set_compute_dtype(torch.float16)
    ofa_b200.set_impl(B.IMPL_AUTO)
    ofa_b200.set_train_dtype(torch.float32)
    yield
    ofa_b200.set_compute_dtype(torch.float16)
    ofa_b200.set_impl(B.IMPL_AUTO)
    ofa_b200.set_train_dtype(torch.float32)


def relerr(a, b):
    a = a.detach().float().cpu().double()
    b = b.detach().float().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / max(1e-12, float(b.abs().max())))


def build(kind, pd, wseed, dev):
    from ofa_b200.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
    cls = OFAMobileNetS4 if kind == 's4' else OFAMobileNetX4
    net = cls(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=list(pd))
    spec = O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], list(pd))
    sd = O.synth_state_dict(spec.param_shapes(), wseed)
    net.load_state_dict(sd)
    return net.to(dev).eval(), spec, sd


def oracle_forward(net, spec, x):
    """The oracle on the net's CURRENT parameters (whatever wrote them)."""
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        return O.supernet_forward(x.cpu(), sd, spec)


# =================================================================================================
# derived 16-bit weight copies can never be stale: every writer the reference uses
# =================================================================================================
def _writers():
    def init_model(net):
        torch.manual_seed(5)
        net.init_model('he_fout')              # m.weight.data.normal_: Tensor._version does not move (ofa/utils.py:134-155)

    def data_copy(net):
        g = torch.Generator(device='cpu').manual_seed(9)
        for name, p in net.named_parameters():
            if name.endswith('conv.weight'):
                p.data.copy_((torch.randn(p.shape, generator=g) * (0.5 / np.sqrt(p[0].numel()))).to(p.device))

    def data_mul(net):
        for name, p in net.named_parameters():
            if 'point_linear.conv' in name or 'inverted_bottleneck.conv' in name:
                p.data.mul_(0.5)

    def reorganize(net):
        net.re_organize_middle_weights()       # dynamic_layers.py:156-199 (permutes through .data)

    def load_sd(net):
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        for k in sd:
            if k.endswith('conv.weight'):
                sd[k] = sd[k] * 0.75
        net.load_state_dict(sd)
    return dict(init_model=init_model, data_copy=data_copy, data_mul=data_mul, reorganize=reorganize, load_sd=load_sd)


@pytest.mark.parametrize('writer', ['init_model', 'data_copy', 'data_mul', 'reorganize', 'load_sd'])
@pytest.mark.parametrize('shape', [(1, 3, 24, 32), (1, 3, 96, 120)])
def test_inference_weight_copies_follow_every_writer(dev, writer, shape):
    """forward -> <writer> -> forward: the second forward must use the NEW weights on the fp16 tensor-core path (NHWC
    kernels for the small image, planar kernels for the 96 x 120 frame) -- compared with the oracle run on the net's
    current parameters.  Also twice in a row with the same input shape, so the second pass takes the recorded PLAN (one
    multi-job pack launch) rather than the first-use packs."""
    import ofa_b200
    net, spec, _ = build('s4', [1, 2], 71, dev)
    sub = dict(ks=5, e=4, d=3, pixel_d=2) if writer != 'reorganize' else dict(ks=7, e=3, d=4, pixel_d=2)
    net.set_active_subnet(**sub)
    spec.set_active_subnet(**sub)
    x = torch.from_numpy(np.random.RandomState(2).rand(*shape).astype(np.float32))
    ofa_b200.set_compute_dtype(torch.float16)
    with torch.no_grad():
        for _ in range(2):                      # second pass = plan replay
            y0 = net(x.to(dev))
        assert relerr(y0, oracle_forward(net, spec, x)) < 1e-2
        _writers()[writer](net)
        ref1 = oracle_forward(net, spec, x)
        assert relerr(ref1, y0.cpu()) > 2e-2 or writer == 'reorganize', 'the writer did not change the function'
        y1 = net(x.to(dev))
        assert relerr(y1, ref1) < 1e-2, (writer, relerr(y1, ref1))
        y2 = net(x.to(dev))
        assert torch.equal(y1, y2)


@pytest.mark.parametrize('writer', ['init_model', 'data_copy', 'data_mul'])
def test_training_weight_copies_follow_every_writer(dev, writer):
    """The bf16 training path (forward packs 'f', data-gradient packs 'b'): step -> <writer> -> step.  The second step's
    loss and the stem-adjacent weight gradient must match the exact fp32 CUDA-core path on the SAME (new) weights."""
    import ofa_b200
    net, spec, _ = build('s4', [2], 72, dev)
    net.train()
    net.set_active_subnet(ks=5, e=4, d=3, pixel_d=2)
    rs = np.random.RandomState(4)
    x = torch.from_numpy(rs.rand(4, 3, 16, 16).astype(np.float32)).to(dev)
    tgt = torch.from_numpy(rs.rand(4, 3, 64, 64).astype(np.float32)).to(dev)

    def step(dtype):
        ofa_b200.set_train_dtype(dtype)
        net.zero_grad(set_to_none=True)
        bn_state = copy.deepcopy({k: v.clone() for k, v in net.state_dict().items() if 'running' in k or 'tracked' in k})
        loss = torch.nn.functional.mse_loss(net(x), tgt)
        loss.backward()
        g = net.blocks[1].mobile_inverted_conv.point_linear.conv.conv.weight.grad.clone()
        net.load_state_dict(bn_state, strict=False)       # keep the BN buffers identical for the twin run
        return float(loss.detach()), g

    step(torch.bfloat16)
    _writers()[writer](net)
    l16, g16 = step(torch.bfloat16)
    l32, g32 = step(torch.float32)
    assert abs(l16 - l32) < 0.02 * abs(l32), (writer, l16, l32)
    cos = float((g16 * g32).sum() / (g16.norm() * g32.norm()))
    assert cos > 0.95, (writer, cos)      # stale weights give ~0; bf16 rounding alone costs 1-2 % here


def test_standalone_module_and_functional_calls_are_never_stale(dev):
    """A DynamicMBConvLayer used on its own (no network-level forward scope) and a bare functional conv call repack on
    every call."""
    from ofa_b200.elastic_nn.modules.dynamic_layers import DynamicMBConvLayer
    from ofa_b200 import functional as OF, backend as B
    torch.manual_seed(0)
    blk = DynamicMBConvLayer([64], [64], [3, 5, 7], [3, 4, 6]).to(dev).eval()
    for m in blk.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_var.uniform_(0.5, 1.5)
            m.running_mean.normal_(0, 0.1)
    x = torch.randn(1, 64, 24, 32, device=dev).half().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y0 = blk(x).float()
        blk.point_linear.conv.conv.weight.data.mul_(2.0)
        y1 = blk(x).float()
        assert relerr(y1, y0) > 0.1
        # bare functional call with a caller-owned cache object
        w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
        cache = OF.PackedWeightCache()
        a = OF.conv_bn_act_infer(x, w, 64, 64, 3, None, B.ACT_NONE, cache=cache).float()
        w.data.mul_(3.0)
        b = OF.conv_bn_act_infer(x, w, 64, 64, 3, None, B.ACT_NONE, cache=cache).float()
        assert relerr(b, 3.0 * a) < 5e-3


def test_cuda_graph_replay_follows_weight_updates(dev):
    """The plan's multi-job pack launch is part of a captured forward: replays re-derive the 16-bit copies from the
    CURRENT fp32 masters (round 1 froze them at capture time)."""
    net, spec, _ = build('s4', [1, 2], 73, dev)
    net.set_active_subnet(ks=3, e=3, d=2, pixel_d=1)
    spec.set_active_subnet(ks=3, e=3, d=2, pixel_d=1)
    x = torch.from_numpy(np.random.RandomState(6).rand(1, 3, 32, 40).astype(np.float32))
    xs = x.to(dev)
    with torch.no_grad():
        for _ in range(3):
            net(xs)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = net(xs)
        g.replay()
        torch.cuda.synchronize()
        assert relerr(y, oracle_forward(net, spec, x)) < 1e-2
        for p in net.parameters():
            if p.dim() == 4:
                p.data.mul_(0.9)
        g.replay()
        torch.cuda.synchronize()
        assert relerr(y, oracle_forward(net, spec, x)) < 1e-2


# =================================================================================================
# ConvLayer(act_func='pixelshuffle') with cout / 4 not a multiple of 32
# =================================================================================================
@pytest.mark.parametrize('cout', [64, 192, 256])
@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
def test_pixelshuffle_conv_any_group_width(dev, cout, dtype):
    from ofa_b200 import functional as OF, backend as B
    rs = np.random.RandomState(cout)
    x = torch.from_numpy(rs.randn(1, 64, 20, 24).astype(np.float32))
    w = torch.from_numpy((rs.randn(cout, 64, 3, 3) * 0.05).astype(np.float32))
    ref = torch.nn.functional.pixel_shuffle(torch.nn.functional.conv2d(x.to(dtype).float(), w.to(dtype).float(), padding=1), 2)
    xd = x.to(dev).to(dtype).contiguous(memory_format=torch.channels_last)
    canary = torch.full((1, cout // 4, 40, 48 + 8), 7.0, dtype=dtype, device=dev)   # guard band right of the output
    y = OF.conv_bn_act_infer(xd, w.to(dev), 64, cout, 3, None, B.ACT_NONE, store=B.STORE_PIXELSHUFFLE2,
                             cache=OF.PackedWeightCache(), out_dtype=dtype)
    assert tuple(y.shape) == (1, cout // 4, 40, 48)
    tol = 2 ** -8 if dtype == torch.float16 else 2 ** -6
    assert relerr(y, ref) < tol
    assert bool((canary == 7.0).all())


# =================================================================================================
# the single-launch, L2-resident band-scheduled block (csrc/mbconv_band.cu) == the three planar kernels, bit for bit
# =================================================================================================
def _block_args(dev, seed, mid):
    g = torch.Generator(device='cpu').manual_seed(seed)
    mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
    w_exp, w_dw, w_proj = mk(384, 64, 1, 1, sc=0.2), mk(384, 1, 7, 7, sc=0.15), mk(64, 384, 1, 1, sc=0.1)
    m75 = (torch.eye(25) + 0.05 * torch.randn(25, 25, generator=g)).to(dev)
    m53 = (torch.eye(9) + 0.05 * torch.randn(9, 9, generator=g)).to(dev)

    class BN:
        def __init__(self, c):
            self.weight = (torch.rand(c, generator=g) + 0.5).to(dev)
            self.bias = (torch.randn(c, generator=g) * 0.1).to(dev)
            self.running_mean = (torch.randn(c, generator=g) * 0.1).to(dev)
            self.running_var = (torch.rand(c, generator=g) + 0.5).to(dev)
            self.eps = 1e-5
    return w_exp, w_dw, m75, m53, w_proj, BN(384), BN(384), BN(64)


@pytest.mark.parametrize('mid,ks,n,h,w', [(384, 7, 1, 200, 120), (192, 3, 1, 96, 120), (256, 5, 2, 140, 344),
                                           (384, 5, 1, 64, 448), (384, 7, 1, 300, 960), (384, 3, 1, 130, 232)])
@pytest.mark.parametrize('res', [True, False])
def test_band_block_is_bit_identical_to_three_kernels(dev, mid, ks, n, h, w, res):
    from ofa_b200 import functional as OF, backend as B
    import ofa_b200
    w_exp, w_dw, m75, m53, w_proj, b1, b2, b3 = _block_args(dev, mid + ks, mid)
    x = torch.randn(n, 64, h, w, device=dev).half().contiguous(memory_format=torch.channels_last)
    out = {}
    for impl in (B.IMPL_PLANAR3, B.IMPL_BAND):
        ofa_b200.set_impl(impl)
        with torch.no_grad():
            out[impl] = OF.mbconv_infer(x, w_exp, w_dw, m75, m53, w_proj, 64, mid, 64, ks, True, B.ACT_RELU6, b1, b2, b3, res,
                                        pack_cache={})
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out[B.IMPL_BAND].float()).all())
    assert torch.equal(out[B.IMPL_BAND], out[B.IMPL_PLANAR3])


# =================================================================================================
# batches of small planes on the planar tcgen05 path (the whole-warp filter build made them viable)
# =================================================================================================
@pytest.mark.parametrize('n,h,w', [(64, 24, 24), (64, 48, 48), (5, 40, 56)])
@pytest.mark.parametrize('ks,e', [(7, 6), (3, 3), (5, 4)])
def test_planar_block_on_batches_of_small_planes(dev, n, h, w, ks, e):
    """64 x 384 planes of 24 x 24 / 48 x 48 pixels (the X4 teacher's / the training shapes), forced onto the planar
    kernels: one M = 64 Toeplitz tile per plane, filter tiles rebuilt for every tile.  Against the oracle block."""
    from ofa_b200.elastic_nn.modules.dynamic_layers import DynamicMBConvLayer
    from ofa_b200.layers import MobileInvertedResidualBlock, IdentityLayer
    from ofa_b200 import backend as B
    import ofa_b200
    torch.manual_seed(ks * 10 + e)
    mb = DynamicMBConvLayer([64], [64], [3, 5, 7], [3, 4, 6])
    blk = MobileInvertedResidualBlock(mb, IdentityLayer(64, 64)).to(dev).eval()
    for m in blk.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_var.uniform_(0.5, 1.5)
            m.running_mean.normal_(0, 0.1)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    for name, p in blk.named_parameters():
        if name.endswith('_matrix'):
            p.data.add_(0.05 * torch.randn_like(p))
    mb.active_kernel_size, mb.active_expand_ratio = ks, e
    x = torch.randn(n, 64, h, w)
    sd = {'b.' + k: v.detach().cpu() for k, v in blk.state_dict().items()}
    with torch.no_grad():
        ref = O.mbconv_block(x, sd, 'b.', ks, e, [3, 5, 7])           # includes the identity shortcut
        ofa_b200.set_impl(B.IMPL_PLANAR3)
        y = blk(x.to(dev).half().contiguous(memory_format=torch.channels_last))
    assert relerr(y, ref) < 6e-3, relerr(y, ref)


# =================================================================================================
# C5 on the kernels the headline runs: sampled per-block (ks, e) mixtures on the planar tcgen05 path
# =================================================================================================
def test_sampled_subnets_on_the_planar_frame_path(dev):
    """50 sampled sub-networks per net on a 96 x 128 LR frame (S4) / 192 x 256 HR frame (X4): IMPL_AUTO takes the planar
    kernels (planes >= 8192 pixels), every block with its own (ks, e).  fp16 storage against the oracle."""
    import ofa_b200
    from ofa_b200 import functional as OF
    ofa_b200.set_compute_dtype(torch.float16)
    for kind, shape in (('s4', (1, 3, 96, 128)), ('x4', (1, 3, 192, 256))):
        net, spec, sd = build(kind, [1, 2], 43, dev)
        x = torch.from_numpy(np.random.RandomState(5).rand(*shape).astype(np.float32))
        probe = torch.empty(1, 64, shape[2] // (1 if kind == 's4' else 2), shape[3] // (1 if kind == 's4' else 2),
                            dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
        assert OF.planar_supported(probe, 64, 384, 64) and OF.planar_preferred(probe)
        worst = 0.0
        for seed in range(50):
            random.seed(1000 + seed)
            a = net.sample_active_subnet()
            random.seed(1000 + seed)
            b = spec.sample_active_subnet()
            assert a == b and list(net.runtime_depth) == spec.runtime_depth
            with torch.no_grad():
                ref = O.supernet_forward(x, sd, spec)
                y = net(x.to(dev))
            worst = max(worst, relerr(y, ref))
            assert relerr(y, ref) < 1e-2, (kind, seed, relerr(y, ref))
        print('planar-path sweep %s: worst max-rel-err %.2e over 50 sub-networks' % (kind, worst))


def test_fp32_sweep_200_seeds_and_structured_grid(dev):
    """SURVEY 8d C5 in full: seeds 0..199 per net on the exact fp32 path plus the 3 x 3 x 3 x 2 grid of the reference's
    `validate` (progressive_shrinking.py:45-59)."""
    import ofa_b200
    ofa_b200.set_compute_dtype(torch.float32)
    for kind, shape in (('s4', (1, 3, 10, 16)), ('x4', (1, 3, 16, 32))):
        net, spec, sd = build(kind, [1, 2], 44, dev)
        x = torch.from_numpy(np.random.RandomState(7).rand(*shape).astype(np.float32))
        for seed in range(100, 200):                        # 0..99 run in test_random_subnet_sweep_vs_oracle
            random.seed(seed)
            a = net.sample_active_subnet()
            random.seed(seed)
            b = spec.sample_active_subnet()
            assert a == b and list(net.runtime_depth) == spec.runtime_depth
            with torch.no_grad():
                assert relerr(net(x.to(dev)), O.supernet_forward(x, sd, spec)) < 1e-3, (kind, seed)
        for d in (2, 3, 4):
            for e in (3, 4, 6):
                for ks in (3, 5, 7):
                    for pd in (1, 2):
                        sub = dict(ks=ks, e=e, d=d, pixel_d=pd)
                        net.set_active_subnet(**sub)
                        spec.set_active_subnet(**sub)
                        with torch.no_grad():
                            assert relerr(net(x.to(dev)), O.supernet_forward(x, sd, spec)) < 1e-3, (kind, sub)


# =================================================================================================
# 16-bit evidence
# =================================================================================================
def test_fp16_storage_range_on_adversarial_weights(dev):
    """fp16 trunk storage on weights built to GROW the unclamped residual trunk: BatchNorm gamma ~ 8 on every project
    BN over 16 blocks.  The fp16 path must stay finite and within 1 % of the oracle as long as the oracle's own trunk
    stays below the fp16 range; `ofa_b200.check_finite` reports the overflow case instead of returning inf silently."""
    import ofa_b200
    net, spec, sd = build('s4', [1, 2], 45, dev)
    sub = dict(ks=7, e=6, d=4, pixel_d=2)
    net.set_active_subnet(**sub)
    spec.set_active_subnet(**sub)
    x = torch.from_numpy(np.random.RandomState(9).rand(1, 3, 96, 128).astype(np.float32))
    for gamma, expect_ok in ((8.0, True), (4000.0, False)):
        sd2 = {k: v.clone() for k, v in sd.items()}
        for k in sd2:
            if k.endswith('point_linear.bn.bn.weight'):
                sd2[k] = torch.full_like(sd2[k], gamma)
        net.load_state_dict(sd2)
        with torch.no_grad():
            ref = O.supernet_forward(x, sd2, spec)
            ofa_b200.set_compute_dtype(torch.float16)
            y16 = net(x.to(dev))
            ofa_b200.set_compute_dtype(torch.bfloat16)
            ybf = net(x.to(dev))
        if expect_ok:
            assert bool(torch.isfinite(y16).all())
            assert relerr(y16, ref) < 1e-2, relerr(y16, ref)
        else:
            # the trunk leaves the fp16 range: fp16 storage overflows (documented limit), bf16 storage does not, and the
            # opt-in overflow policy recomputes the frame in bf16
            assert ofa_b200.check_finite(ybf) and not ofa_b200.check_finite(y16)
            ofa_b200.set_compute_dtype(torch.float16)
            ofa_b200.set_overflow_policy('fallback_bf16')
            try:
                with torch.no_grad():
                    y = net(x.to(dev))
            finally:
                ofa_b200.set_overflow_policy('none')
            assert ofa_b200.check_finite(y) and torch.equal(y, ybf)


def test_bf16_training_trajectory_tracks_fp32(dev):
    """30 optimizer steps of the same sampled sub-network sequence on the same data: bf16 mixed precision against the
    exact fp32 path.  Loss curves within 3 % at every step; PSNR-Y of a held-out pair within 0.15 dB at the end (the nets
    are 30 steps away from a random initialisation: ~10 dB, where a 1 % change of the MSE is already 0.04 dB, and the
    weight-gradient reductions use atomics, so the runs differ from launch to launch by ~0.05 dB)."""
    import ofa_b200
    from ofa_b200 import optim, metrics
    from ofa_b200.elastic_nn.training import train_step
    rs = np.random.RandomState(12)
    hr = torch.from_numpy(rs.rand(8, 3, 64, 64).astype(np.float32)).to(dev)
    lr_img = torch.nn.functional.avg_pool2d(hr, 4)
    held_hr = torch.from_numpy(rs.rand(2, 3, 64, 64).astype(np.float32)).to(dev)
    held_lr = torch.nn.functional.avg_pool2d(held_hr, 4)
    batch = {'image': hr, '2x_down_image': torch.nn.functional.avg_pool2d(hr, 2), '4x_down_image': lr_img}
    curves, finals = {}, {}
    for name, dtype in (('fp32', torch.float32), ('bf16', torch.bfloat16)):
        ofa_b200.set_train_dtype(dtype)
        net, _, _ = build('s4', [2], 46, dev)
        net.train()
        decay, no_decay = optim.split_no_decay(net.named_parameters())
        opt = optim.FusedAdam(decay, no_decay, lr=1e-3, weight_decay=3e-5)
        losses = []
        for i in range(30):
            loss, _, _ = train_step(net, opt, batch, 0, i, 1000, dynamic_batch_size=2)
            losses.append(float(loss))
        curves[name] = losses
        net.eval()
        net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
        ofa_b200.set_compute_dtype(torch.float32)
        with torch.no_grad():
            finals[name] = metrics.psnr_y(net(held_lr), held_hr)
    worst = max(abs(a - b) / abs(a) for a, b in zip(curves['fp32'], curves['bf16']))
    print('trajectory: fp32 %.4f -> %.4f, bf16 %.4f -> %.4f, worst relative gap %.3f; held-out PSNR-Y %.3f / %.3f dB' % (
        curves['fp32'][0], curves['fp32'][-1], curves['bf16'][0], curves['bf16'][-1], worst, finals['fp32'], finals['bf16']))
    assert curves['fp32'][-1] < 0.8 * curves['fp32'][0], 'the fp32 run did not train'
    assert worst < 0.03, worst
    assert abs(finals['fp32'] - finals['bf16']) < 0.15, finals


def test_fused_adam_state_dict_round_trip(dev):
    """optimizer.state_dict() / load_state_dict (sr_run_manager.py:301,539) and param_group['lr'] writes (:78-90)."""
    from ofa_b200 import optim
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(300, device=dev)), torch.nn.Parameter(torch.randn(17, 5, device=dev))]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a = optim.FusedAdam([ps[0]], [ps[1]], lr=1e-2, weight_decay=1e-2)
    t = torch.optim.Adam([{'params': [ref[0]], 'weight_decay': 1e-2}, {'params': [ref[1]], 'weight_decay': 0.0}], lr=1e-2)

    def run(opt, params, steps, seed0):
        for s in range(steps):
            g = torch.Generator(device='cpu').manual_seed(seed0 + s)
            for p in params:
                p.grad = torch.randn(p.shape, generator=g).to(dev)
            opt.step()
    run(a, ps, 3, 0)
    run(t, ref, 3, 0)
    saved = a.state_dict()
    b = optim.FusedAdam([ps[0]], [ps[1]], lr=5e-1, weight_decay=1e-2)        # fresh optimizer, wrong lr
    b.load_state_dict(saved)
    assert b.param_groups[0]['lr'] == 1e-2
    for g in b.param_groups:
        g['lr'] = 5e-3                                                       # the run manager's write
    for g in t.param_groups:
        g['lr'] = 5e-3
    run(b, ps, 3, 10)
    run(t, ref, 3, 10)
    for p, r in zip(ps, ref):
        assert relerr(p, r) < 1e-5


# =================================================================================================
# image output formats of the inference path
# =================================================================================================
@pytest.mark.parametrize('kind,shape', [('s4', (2, 3, 40, 56)), ('s4', (1, 3, 96, 120)), ('x4', (1, 3, 64, 96))])
def test_uint8_and_fp16_image_outputs(dev, kind, shape):
    """set_output_dtype(torch.uint8): the last conv's epilogue writes round_half_even(clamp(y, 0, 1) * 255) -- exactly
    the reference's tensor2img_np (sr_run_manager.py:567-597) applied to the fp32 output of the same forward, bit for bit;
    torch.float16: the fp32 values rounded once."""
    net, spec, sd = build(kind, [1, 2], 47, dev)
    sub = dict(ks=5, e=4, d=3, pixel_d=2)
    net.set_active_subnet(**sub)
    x = torch.from_numpy(np.random.RandomState(3).rand(*shape).astype(np.float32)).to(dev)
    with torch.no_grad():
        y32 = net(x)
        # put the random-weight output into the image range so that the clamp and every uint8 level are exercised
        lo, hi = float(y32.min()), float(y32.max())
        last = net.dec_final_output_conv_block
        last.bn.weight.data.mul_(1.2 / (hi - lo))
        last.bn.bias.data.copy_((last.bn.bias.data - lo) * (1.2 / (hi - lo)) - 0.1)
        last.bn.running_mean.data.mul_(1.0)
        y32 = net(x)
        net.set_output_dtype(torch.uint8)
        y8 = net(x)
        net.set_output_dtype(torch.float16)
        y16 = net(x)
        net.set_output_dtype(torch.float32)
    assert y8.dtype == torch.uint8 and y8.shape == y32.shape and y8.is_contiguous()
    want = (y32.clamp(0, 1) * 255.0).round().to(torch.uint8)          # torch.round = half to even, as numpy's
    assert torch.equal(y8, want)
    assert int(want.min()) == 0 and int(want.max()) == 255 and len(torch.unique(want)) > 200
    assert y16.dtype == torch.float16 and torch.equal(y16, y32.to(torch.float16))
    # the reference metric on the uint8 image equals the metric computed from the fp32 tensor
    t = torch.rand_like(y32)
    from ofa_b200 import metrics
    assert metrics.psnr_y(y32, t) == pytest.approx(metrics.psnr_y(y8.float() / 255.0, t), abs=1e-9)


def test_graphed_module_matches_eager_and_follows_subnet_and_weights(dev):
    """ofa_b200.GraphedModule: bit-identical to the eager forward, one graph per (shape, sub-network), replays see
    weight updates, `copies=2` alternates output buffers."""
    import ofa_b200
    net, spec, _ = build('s4', [1, 2], 48, dev)
    fast = ofa_b200.GraphedModule(net, copies=2)
    x = torch.from_numpy(np.random.RandomState(4).rand(1, 3, 40, 48).astype(np.float32)).to(dev)
    with torch.no_grad():
        for sub in (dict(ks=7, e=6, d=4, pixel_d=2), dict(ks=3, e=3, d=2, pixel_d=1)):
            net.set_active_subnet(**sub)
            y_eager = net(x)
            y1 = fast(x).clone()
            y2 = fast(x)                       # second copy
            assert torch.equal(y1, y_eager) and torch.equal(y2, y_eager)
        for p in net.parameters():
            if p.dim() == 4:
                p.data.mul_(0.9)
        y_eager = net(x)
        assert torch.equal(fast(x), y_eager) and torch.equal(fast(x), y_eager)
    assert len(fast._graphs) == 4
