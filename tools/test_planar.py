"""GPU bring-up check of the three planar MBConv stages against torch ops on the same device, plus
timings at the bench shape.  (Debug tool; the parity tests proper live in tests/ and use the oracle.)
    python tools/test_planar.py [dw|expand|project|block|time] ...
"""
import os
import sys
from ctypes import byref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ofa-for-super-resolution_b200'))
import torch
import torch.nn.functional as F
from ofa_b200 import backend as B

which = sys.argv[1:] or ['dw', 'expand', 'project', 'time']
dev = torch.device('cuda:0')
L = B.lib()
st = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)


class BN:
    def __init__(self, c):
        self.weight = (torch.rand(c, device=dev) + 0.5)
        self.bias = torch.randn(c, device=dev) * 0.1
        self.running_mean = torch.randn(c, device=dev) * 0.1
        self.running_var = torch.rand(c, device=dev) + 0.5
        self.eps = 1e-5

    def s(self):
        return B.OfaBn(self.weight.data_ptr(), self.bias.data_ptr(), self.running_mean.data_ptr(),
                       self.running_var.data_ptr(), self.eps)

    def fold(self, c):
        sc = self.weight[:c] / torch.sqrt(self.running_var[:c] + self.eps)
        return sc, self.bias[:c] - self.running_mean[:c] * sc


def tdt(code):
    return torch.float16 if code == B.OFA_F16 else torch.bfloat16


def report(name, got, ref, tol):
    err = float((got.float() - ref.float()).abs().max())
    scale = float(ref.float().abs().max())
    ok = err <= tol * max(scale, 1e-6)
    print('%-44s max|err| %.4e  max|ref| %.3f  %s' % (name, err, scale, 'OK' if ok else 'MISMATCH'), flush=True)
    return ok


def active_filter(w7, m75, m53, ks):
    C = w7.shape[0]
    if ks == 7:
        return w7
    k5 = F.linear(w7[:, :, 1:6, 1:6].reshape(C, 25), m75).view(C, 1, 5, 5)
    if ks == 5:
        return k5
    return F.linear(k5[:, :, 1:4, 1:4].reshape(C, 9), m53).view(C, 1, 3, 3)


def run_dw(dtype, ks, N, C, H, W, timing=False):
    dt = tdt(dtype)
    x = (torch.rand(N, C, H, W, device=dev) * 6).to(dt)
    w7 = torch.randn(C, 1, 7, 7, device=dev) * 0.15
    m75 = torch.eye(25, device=dev) + 0.05 * torch.randn(25, 25, device=dev)
    m53 = torch.eye(9, device=dev) + 0.05 * torch.randn(9, 9, device=dev)
    bn = BN(C)
    y = torch.full((N, C, H, W), 7.0, device=dev).to(dt)
    bs = bn.s()

    def call():
        B.check(L.ofa_dw_planar_fwd(x.data_ptr(), y.data_ptr(), N, C, H, W, w7.data_ptr(), 7, m75.data_ptr(),
                                    m53.data_ptr(), 1, ks, dtype, byref(bs), B.ACT_RELU6, st()))
    call()
    torch.cuda.synchronize()
    if timing:
        return call
    f = active_filter(w7, m75, m53, ks).to(dt).float()
    ref = F.conv2d(x.float(), f, padding=ks // 2, groups=C)
    sc, sh = bn.fold(C)
    ref = torch.clamp(ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1), 0, 6)
    return report('dw ks=%d %s N%d C%d %dx%d' % (ks, str(dt)[6:], N, C, H, W), y, ref, 6e-3 if dtype == B.OFA_F16 else 1.2e-2)


def pack(w_exp, w_proj, mid, dtype, trunk=None):
    trunk = trunk or B.OFA_BF16
    mt = (mid + 127) // 128
    we = torch.empty(mt * 128, 64, dtype=tdt(trunk), device=dev)
    wp = torch.empty(64, mid, dtype=tdt(dtype), device=dev)
    B.check(L.ofa_mbconv_pack_weights(w_exp.data_ptr(), w_exp.stride(0), w_exp.stride(1), w_proj.data_ptr(),
                                      w_proj.stride(0), w_proj.stride(1), mid, trunk, dtype, we.data_ptr(), wp.data_ptr(),
                                      st()))
    return we, wp


def run_expand(dtype, mid, N, HW, timing=False, trunk=B.OFA_BF16):
    dt = tdt(dtype)
    x = torch.randn(N, HW, 64, device=dev).to(tdt(trunk))
    w_exp = torch.randn(384, 64, 1, 1, device=dev) * 0.2
    w_proj = torch.randn(64, 384, 1, 1, device=dev) * 0.1
    we, _ = pack(w_exp, w_proj, mid, dtype, trunk)
    bn = BN(384)
    y = torch.full((N, mid, HW), 7.0, device=dev).to(dt)
    bs = bn.s()

    def call():
        B.check(L.ofa_expand_planar_fwd(x.data_ptr(), y.data_ptr(), we.data_ptr(), N, HW, mid, trunk, dtype, byref(bs),
                                        B.ACT_RELU6, st()))
    call()
    torch.cuda.synchronize()
    if timing:
        return call
    wr = w_exp[:mid, :, 0, 0].to(tdt(trunk)).float()
    ref = torch.einsum('npk,mk->nmp', x.float(), wr)
    sc, sh = bn.fold(mid)
    ref = torch.clamp(ref * sc.view(1, -1, 1) + sh.view(1, -1, 1), 0, 6)
    return report('expand mid=%d %s trunk %s N%d HW%d' % (mid, str(dt)[6:], str(tdt(trunk))[6:], N, HW), y, ref, 3e-3 if dtype == B.OFA_F16 else 1e-2)


def run_project(dtype, mid, N, HW, res=True, timing=False, trunk=B.OFA_BF16):
    dt = tdt(dtype)
    x = (torch.rand(N, mid, HW, device=dev) * 6).to(dt)
    w_exp = torch.randn(384, 64, 1, 1, device=dev) * 0.2
    w_proj = torch.randn(64, 384, 1, 1, device=dev) * 0.1
    _, wp = pack(w_exp, w_proj, mid, dtype, trunk)
    bn = BN(64)
    r = torch.randn(N, HW, 64, device=dev).to(tdt(trunk))
    y = torch.full((N, HW, 64), 7.0, device=dev).to(tdt(trunk))
    bs = bn.s()

    def call():
        B.check(L.ofa_project_planar_fwd(x.data_ptr(), r.data_ptr() if res else None, y.data_ptr(), wp.data_ptr(), N,
                                         HW, mid, trunk, dtype, byref(bs), st()))
    call()
    torch.cuda.synchronize()
    if timing:
        return call
    wr = w_proj[:, :mid, 0, 0].to(dt).float()
    ref = torch.einsum('nkp,ok->npo', x.float(), wr)
    sc, sh = bn.fold(64)
    ref = ref * sc.view(1, 1, -1) + sh.view(1, 1, -1)
    if res:
        ref = ref + r.float()
    return report('project mid=%d %s trunk %s N%d HW%d res=%d' % (mid, str(dt)[6:], str(tdt(trunk))[6:], N, HW, res), y, ref,
                  1e-2 if trunk == B.OFA_BF16 else 2e-3)


def timeit(name, call, nbytes, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        call()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print('%-28s %.3f ms   %.0f GB/s algorithmic' % (name, ms, nbytes / ms / 1e6), flush=True)


ok = True
if 'dw' in which:
    for dtype in (B.OFA_F16, B.OFA_BF16):
        for ks in (7, 5, 3):
            ok &= run_dw(dtype, ks, 1, 3, 200, 120)
    ok &= run_dw(B.OFA_F16, 7, 2, 5, 40, 56)
    ok &= run_dw(B.OFA_F16, 3, 1, 300, 130, 64)
if 'expand' in which:
    for dtype in (B.OFA_F16, B.OFA_BF16):
        for mid in (384, 256, 192):
            ok &= run_expand(dtype, mid, 1, 1000)
    ok &= run_expand(B.OFA_F16, 384, 3, 40 * 56)
    ok &= run_expand(B.OFA_F16, 192, 1, 540 * 960)
    ok &= run_expand(B.OFA_F16, 384, 2, 1000, trunk=B.OFA_F16)
if 'project' in which:
    for dtype in (B.OFA_F16, B.OFA_BF16):
        for mid in (384, 256, 192):
            ok &= run_project(dtype, mid, 1, 1000)
    ok &= run_project(B.OFA_F16, 384, 3, 40 * 56, res=False)
    ok &= run_project(B.OFA_F16, 192, 1, 540 * 960)
    ok &= run_project(B.OFA_F16, 384, 2, 1000, trunk=B.OFA_F16)
if 'time' in which:
    H, W = 540, 960
    P = H * W
    timeit('expand 64->384', run_expand(B.OFA_F16, 384, 1, P, timing=True), P * (64 + 384) * 2)
    timeit('dw7 C384', run_dw(B.OFA_F16, 7, 1, 384, H, W, timing=True), 2 * P * 384 * 2)
    timeit('dw5 C384', run_dw(B.OFA_F16, 5, 1, 384, H, W, timing=True), 2 * P * 384 * 2)
    timeit('dw3 C192', run_dw(B.OFA_F16, 3, 1, 192, H, W, timing=True), 2 * P * 192 * 2)
    timeit('project 384->64 +res', run_project(B.OFA_F16, 384, 1, P, timing=True), P * (384 + 128) * 2)
def run_block(impl, mid, ks, N, H, W, res=True, dtype=None, seed=0, timing=False, trunk=None):
    """Whole block through ofa_mbconv_fwd with the given impl (IMPL_BAND: one launch; IMPL_PLANAR3: three kernels)."""
    from ofa_b200 import functional as OF
    dtype = dtype or B.OFA_F16
    trunk = trunk or B.OFA_F16
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(N, 64, H, W, device=dev, generator=g).to(tdt(trunk)).contiguous(memory_format=torch.channels_last)
    w_exp = torch.randn(384, 64, 1, 1, device=dev, generator=g) * 0.2
    w_dw = torch.randn(384, 1, 7, 7, device=dev, generator=g) * 0.15
    w_proj = torch.randn(64, 384, 1, 1, device=dev, generator=g) * 0.1
    m75 = torch.eye(25, device=dev) + 0.05 * torch.randn(25, 25, device=dev, generator=g)
    m53 = torch.eye(9, device=dev) + 0.05 * torch.randn(9, 9, device=dev, generator=g)
    torch.manual_seed(seed + 1)
    b1, b2, b3 = BN(384), BN(384), BN(64)

    class _B:
        pass
    bns = []
    for b in (b1, b2, b3):
        o = _B(); o.weight, o.bias, o.running_mean, o.running_var, o.eps = b.weight, b.bias, b.running_mean, b.running_var, b.eps
        bns.append(o)
    OF.set_impl(impl)
    OF.set_mid_dtype(tdt(dtype))

    def call():
        return OF.mbconv_infer(x, w_exp, w_dw, m75, m53, w_proj, 64, mid, 64, ks, True, B.ACT_RELU6, bns[0], bns[1], bns[2],
                               res, pack_cache={})
    y = call()
    torch.cuda.synchronize()
    OF.set_impl(B.IMPL_AUTO)
    if timing:
        def again():
            OF.set_impl(impl)
            r = call()
            OF.set_impl(B.IMPL_AUTO)
            return r
        return again
    return y


if 'band' in which:
    shapes = [(384, 7, 1, 200, 120), (384, 7, 1, 130, 232), (192, 3, 1, 96, 120), (256, 5, 2, 140, 344),
              (384, 5, 1, 64, 448), (384, 7, 1, 300, 960), (384, 7, 1, 540, 960), (192, 3, 1, 540, 960)]
    for mid, ks, N, H, W in shapes:
        for res in (True, False):
            ref = run_block(B.IMPL_PLANAR3, mid, ks, N, H, W, res)
            got = run_block(B.IMPL_BAND, mid, ks, N, H, W, res)
            bad = int((got.float() != ref.float()).sum())
            err = float((got.float() - ref.float()).abs().max())
            fin = bool(torch.isfinite(got.float()).all())
            print('band vs 3-kernel  mid %d ks %d N%d %dx%d res %d : %d differing elements, max|diff| %.3e, finite %s  %s'
                  % (mid, ks, N, H, W, res, bad, err, fin, 'OK' if bad == 0 and fin else 'MISMATCH'), flush=True)
            ok &= (bad == 0 and fin)
if 'bandtime' in which:
    H, W = 540, 960
    P = H * W
    for mid, ks in ((384, 7), (384, 5), (256, 5), (192, 3)):
        timeit('3-kernel block M%d ks%d' % (mid, ks), run_block(B.IMPL_PLANAR3, mid, ks, 1, H, W, timing=True), 2 * P * 64 * 2)
        timeit('band block     M%d ks%d' % (mid, ks), run_block(B.IMPL_BAND, mid, ks, 1, H, W, timing=True), 2 * P * 64 * 2)
if 'small' in which:
    # batches of small planes (the X4 teacher / training shapes): planar tcgen05 path against the three NHWC kernels
    for (N, H, W) in ((64, 48, 48), (64, 24, 24), (16, 96, 96), (8, 64, 64), (1, 96, 120)):
        for mid, ks in ((384, 7), (384, 3), (192, 5)):
            P = N * H * W
            ref = run_block(B.IMPL_NHWC, mid, ks, N, H, W, True)
            got = run_block(B.IMPL_PLANAR3, mid, ks, N, H, W, True)
            err = float((got.float() - ref.float()).abs().max() / ref.float().abs().max())
            print('N%d %dx%d M%d ks%d  planar vs NHWC max-rel-diff %.2e' % (N, H, W, mid, ks, err), flush=True)
            timeit('  planar3 N%d %dx%d M%d ks%d' % (N, H, W, mid, ks), run_block(B.IMPL_PLANAR3, mid, ks, N, H, W, timing=True), 2 * P * 64 * 2)
            timeit('  nhwc    N%d %dx%d M%d ks%d' % (N, H, W, mid, ks), run_block(B.IMPL_NHWC, mid, ks, N, H, W, timing=True), 2 * P * 64 * 2)
print('ALL OK' if ok else 'SOME MISMATCH')
