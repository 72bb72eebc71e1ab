"""Guard bands around every activation tensor the library writes (compute-sanitizer is not available on the GPU pool).

The module-level allocators of the operator layer (`backend.new_nhwc`, `functional._conv_out`) are replaced by versions
that place each output in the middle of a larger byte buffer pre-filled with a canary; after network forwards and training
steps on RAGGED shapes (tile tails in every direction, widths that are not multiples of 8, batches > 1) every canary byte
in front of and behind every output must be untouched.  This covers the stores of all forward kernels (TMA tensor stores
of the planar / project / thin-output kernels, the vector stores of conv_tc / conv_stem_tc / dw_fast, the BatchNorm and
elementwise kernels) and of the backward chain (data gradients, BatchNorm gradients)."""
import random

import numpy as np
import pytest
import torch

import ofa_sr_oracle as O

pytestmark = pytest.mark.gpu
GUARD = 8192
CANARY = 0xA5
FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])


class _Guards:
    def __init__(self):
        self.bufs = []

    def alloc(self, shape, dtype, device, channels_last):
        n = int(np.prod(shape))
        es = torch.empty((), dtype=dtype).element_size()
        nbytes = n * es
        padded = (nbytes + 255) // 256 * 256
        buf = torch.full((GUARD + padded + GUARD,), CANARY, dtype=torch.uint8, device=device)
        flat = buf[GUARD:GUARD + nbytes].view(dtype)
        self.bufs.append((buf, nbytes))
        if channels_last and len(shape) == 4:
            N, C, H, W = shape
            return flat.view(N, H, W, C).permute(0, 3, 1, 2)
        return flat.view(shape)

    def check(self, what):
        assert self.bufs, what + ': no guarded allocation was made'
        for k, (buf, nbytes) in enumerate(self.bufs):
            front, back = buf[:GUARD], buf[GUARD + nbytes:]
            assert bool((front == CANARY).all()), '%s: write in FRONT of output #%d (%d bytes)' % (what, k, nbytes)
            assert bool((back == CANARY).all()), '%s: write BEHIND output #%d (%d bytes)' % (what, k, nbytes)
        n = len(self.bufs)
        self.bufs = []
        return n


@pytest.fixture()
def guards(monkeypatch):
    from ofa_b200 import backend as B, functional as OF
    g = _Guards()

    def new_nhwc(n, c, h, w, dtype, device):
        return g.alloc((n, c, h, w), dtype, device, True)

    def conv_out(x, cout, store, dtype, nchw=False):
        n, _, h, w = x.shape
        if store == B.STORE_PIXELSHUFFLE2:
            shape = (n, cout // 4, 2 * h, 2 * w)
        elif store == B.STORE_PIXELUNSHUFFLE2:
            shape = (n, cout * 4, h // 2, w // 2)
        else:
            shape = (n, cout, h, w)
        return g.alloc(shape, dtype, x.device, not nchw)

    monkeypatch.setattr(B, 'new_nhwc', new_nhwc)
    monkeypatch.setattr(OF, '_conv_out', conv_out)
    return g


def _net(kind, dev):
    from ofa_b200.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
    from ofa_b200.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
    DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
    cls = OFAMobileNetS4 if kind == 's4' else OFAMobileNetX4
    net = cls(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=[1, 2])
    spec = O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2])
    net.load_state_dict(O.synth_state_dict(spec.param_shapes(), 9))
    return net.to(dev)


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16, torch.float32])
def test_inference_outputs_stay_inside_their_tensors(guards, dtype):
    """S4 / X4 inference on ragged frames: planar tcgen05 path (tile tails in rows and columns, M = 64 tail tiles), the
    column-split path (width % 8 != 0), the NHWC kernels (small planes, batch 3) and the exact fp32 kernels."""
    import ofa_b200
    dev = torch.device('cuda:0')
    ofa_b200.set_compute_dtype(dtype)
    try:
        net = _net('s4', dev).eval()
        rs = np.random.RandomState(3)
        cases = [((1, 3, 130, 232), dict(ks=7, e=6, d=4, pixel_d=2)), ((2, 3, 100, 136), dict(ks=5, e=4, d=3, pixel_d=1)),
                 ((1, 3, 67, 121), dict(ks=3, e=3, d=2, pixel_d=2)), ((3, 3, 24, 24), dict(ks=7, e=6, d=4, pixel_d=2)),
                 ((1, 3, 49, 56), dict(ks=5, e=6, d=2, pixel_d=1))]
        with torch.no_grad():
            for shape, sub in cases:
                net.set_active_subnet(**sub)
                y = net(torch.from_numpy(rs.rand(*shape).astype(np.float32)).to(dev))
                torch.cuda.synchronize()
                assert torch.isfinite(y.float()).all()
                assert guards.check('S4 %s %s %s' % (shape, sub, dtype)) >= 5
            netx = _net('x4', dev).eval()
            for shape, sub in [((1, 3, 136, 200), dict(ks=7, e=6, d=4, pixel_d=2)), ((2, 3, 96, 96), dict(ks=3, e=4, d=2, pixel_d=1))]:
                netx.set_active_subnet(**sub)
                y = netx(torch.from_numpy(rs.rand(*shape).astype(np.float32)).to(dev))
                torch.cuda.synchronize()
                assert torch.isfinite(y.float()).all()
                assert guards.check('X4 %s %s %s' % (shape, sub, dtype)) >= 5
    finally:
        ofa_b200.set_compute_dtype(torch.float16)


@pytest.mark.parametrize('tdt', [torch.bfloat16, torch.float32])
def test_training_step_outputs_stay_inside_their_tensors(guards, tdt):
    """Forward + backward of sampled sub-networks on ragged patches (20 x 28 LR, batch 3): the block-level training calls
    (tensor-core convs, weight gradients, depthwise kernels, BatchNorm) and the exact fp32 path."""
    import ofa_b200
    dev = torch.device('cuda:0')
    ofa_b200.set_train_dtype(tdt)
    try:
        net = _net('s4', dev).train()
        rs = np.random.RandomState(4)
        x = torch.from_numpy(rs.rand(3, 3, 20, 28).astype(np.float32)).to(dev)
        tgt = torch.from_numpy(rs.rand(3, 3, 80, 112).astype(np.float32)).to(dev)
        for step in range(3):
            random.seed(40 + step)
            net.sample_active_subnet()
            net.set_active_subnet(pixel_d=2)
            net.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(net(x), tgt)
            loss.backward()
            torch.cuda.synchronize()
            assert np.isfinite(float(loss.detach()))
            assert guards.check('training step %d %s' % (step, tdt)) >= 20
    finally:
        ofa_b200.set_train_dtype(torch.float32)
