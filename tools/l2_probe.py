"""How fast do the planar MBConv stages run when their intermediates stay in the 126 MB L2?
Back-to-back launches on working sets of growing size, no L2 flush in between (round-2 design probe for the
band-scheduled block: DESIGN.md section 3.7).
    python tools/l2_probe.py
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [sys.argv[0], 'none']
import torch
spec = importlib.util.spec_from_file_location('tp', os.path.join(ROOT, 'tools', 'test_planar.py'))
tp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tp)
B = tp.B
dev = torch.device('cuda:0')


def timed(call, reps=20, warm=3):
    for _ in range(warm):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print('--- torch copy, src+dst working set (MB) -> GB/s (read+write)')
for mb in (4, 8, 16, 24, 32, 48, 64, 96, 128, 256, 1024):
    n = mb * (1 << 20) // 2
    a = torch.empty(n // 2, dtype=torch.float16, device=dev).normal_()
    b = torch.empty_like(a)
    ms = timed(lambda: b.copy_(a), reps=50)
    print('copy  ws %5d MB  %.4f ms  %7.0f GB/s' % (mb, ms, 2 * a.numel() * 2 / ms / 1e6), flush=True)
for mb in (8, 16, 32, 48, 64, 96, 256):
    a = torch.empty(mb * (1 << 20) // 2, dtype=torch.float16, device=dev)
    ms = timed(lambda: a.zero_(), reps=50)
    print('fill  ws %5d MB  %.4f ms  %7.0f GB/s' % (mb, ms, a.numel() * 2 / ms / 1e6), flush=True)

H, W = 540, 960
print('--- dw7 on C planes of 540x960, back to back (in+out MB) ')
for C in (8, 16, 24, 32, 48, 64, 384):
    call = tp.run_dw(B.OFA_F16, 7, 1, C, H, W, timing=True)
    ms = timed(call)
    outs = C * H * W
    print('dw7 C%3d  ws %6.1f MB  %.4f ms  %6.0f GB/s alg  %.2f outputs/clk/SM' % (
        C, 2 * outs * 2 / 1e6, ms, 2 * outs * 2 / ms / 1e6, outs / (ms * 1e-3 * 1.965e9 * 148)), flush=True)
print('--- dw7 on 384 planes of 128 x Wr (one row band of width Wr)')
for Wr in (112, 224, 448, 896):
    call = tp.run_dw(B.OFA_F16, 7, 1, 384, 128, Wr, timing=True)
    ms = timed(call)
    outs = 384 * 128 * Wr
    print('dw7 128x%3d  ws %6.1f MB  %.4f ms  %6.0f GB/s alg  %.2f outputs/clk/SM' % (
        Wr, 2 * outs * 2 / 1e6, ms, 2 * outs * 2 / ms / 1e6, outs / (ms * 1e-3 * 1.965e9 * 148)), flush=True)
print('--- expand / project on P pixels')
for P in (128 * 112, 128 * 224, 128 * 448, 128 * 896, H * W):
    ce = tp.run_expand(B.OFA_F16, 384, 1, P, timing=True, trunk=B.OFA_F16)
    ms = timed(ce)
    print('expand  P %7d  ws %6.1f MB  %.4f ms  %6.0f GB/s alg' % (P, P * 448 * 2 / 1e6, ms, P * 448 * 2 / ms / 1e6), flush=True)
    cp = tp.run_project(B.OFA_F16, 384, 1, P, timing=True, trunk=B.OFA_F16)
    ms = timed(cp)
    print('project P %7d  ws %6.1f MB  %.4f ms  %6.0f GB/s alg' % (P, P * 512 * 2 / 1e6, ms, P * 512 * 2 / ms / 1e6), flush=True)
print('--- chained expand -> dw7 -> project on one region, shared t1 / t2 (stream-ordered launches)')
from ctypes import byref
L = B.lib()
st = lambda: torch.cuda.current_stream().cuda_stream
for Hr, Wr in ((128, 112), (128, 224), (128, 448), (64, 960), (128, 896), (540, 960)):
    P = Hr * Wr
    x = torch.randn(1, P, 64, device=dev).half()
    y = torch.empty_like(x)
    t1 = torch.empty(1, 384, P, device=dev, dtype=torch.float16)
    t2 = torch.empty_like(t1)
    w_exp = torch.randn(384, 64, 1, 1, device=dev) * 0.2
    w_proj = torch.randn(64, 384, 1, 1, device=dev) * 0.1
    we, wp = tp.pack(w_exp, w_proj, 384, B.OFA_F16, B.OFA_F16)
    w7 = torch.randn(384, 1, 7, 7, device=dev) * 0.15
    b1, b2, b3 = tp.BN(384), tp.BN(384), tp.BN(64)
    s1, s2, s3 = b1.s(), b2.s(), b3.s()

    def chain():
        B.check(L.ofa_expand_planar_fwd(x.data_ptr(), t1.data_ptr(), we.data_ptr(), 1, P, 384, B.OFA_F16, B.OFA_F16,
                                        byref(s1), B.ACT_RELU6, st()))
        B.check(L.ofa_dw_planar_fwd(t1.data_ptr(), t2.data_ptr(), 1, 384, Hr, Wr, w7.data_ptr(), 7, None, None, 0, 7,
                                    B.OFA_F16, byref(s2), B.ACT_RELU6, st()))
        B.check(L.ofa_project_planar_fwd(t2.data_ptr(), x.data_ptr(), y.data_ptr(), wp.data_ptr(), 1, P, 384, B.OFA_F16,
                                         B.OFA_F16, byref(s3), st()))
    ms = timed(chain)
    print('chain %3dx%3d  P %7d  t1+t2 %6.1f MB  %.4f ms  = %.3f ms per 540x960 frame-equivalent' % (
        Hr, Wr, P, 2 * P * 768 / 1e6, ms, ms * H * W / P), flush=True)
