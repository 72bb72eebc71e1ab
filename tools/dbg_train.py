import os, sys
sys.path.insert(0, '/root/repo/ofa-for-super-resolution_b200')
import torch
import ofa_b200
from ofa_b200 import functional as OF, backend as B
dev = torch.device('cuda:0')
torch.manual_seed(0)
def cos(a, b):
    a = a.float().flatten(); b = b.float().flatten()
    return float(a @ b / (a.norm() * b.norm()))
for (cin, cout, ks, cmi, cmo) in [(64, 384, 1, 64, 384), (384, 64, 1, 384, 64), (192, 64, 1, 384, 64), (64, 64, 5, 64, 64), (64, 256, 5, 64, 256), (64, 3, 5, 64, 3), (64, 192, 1, 64, 384), (3, 64, 5, 3, 64), (64, 3, 3, 64, 3), (3, 64, 3, 3, 64)]:
    w = (torch.randn(cmo, cmi, ks, ks, device=dev) * 0.1).requires_grad_(True)
    x32 = torch.randn(2, cin, 12, 20, device=dev)
    res = {}
    for dt in (torch.float32, torch.bfloat16):
        ofa_b200.set_train_dtype(dt)
        x = (x32.to(dt) if cin >= 16 else x32.clone()).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = OF.conv2d(x, w, cin, cout, ks)
        g = torch.randn(y.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1)).to(y.dtype).contiguous(memory_format=torch.channels_last)
        w.grad = None
        y.backward(g)
        res[dt] = (y.detach().float(), x.grad.detach().float(), w.grad.detach().float().clone())
    a, b = res[torch.float32], res[torch.bfloat16]
    print('conv %d->%d k%d  y cos %.6f  dx cos %.6f  dw cos %.6f' % (cin, cout, ks, cos(a[0], b[0]), cos(a[1], b[1]), cos(a[2], b[2])), flush=True)
# depthwise
for ks in (3, 5, 7):
    C = 192
    w7 = (torch.randn(384, 1, 7, 7, device=dev) * 0.2).requires_grad_(True)
    m75 = (torch.eye(25, device=dev) + 0.05 * torch.randn(25, 25, device=dev)).requires_grad_(True)
    m53 = (torch.eye(9, device=dev) + 0.05 * torch.randn(9, 9, device=dev)).requires_grad_(True)
    x32 = torch.randn(2, C, 12, 20, device=dev)
    res = {}
    for dt in (torch.float32, torch.bfloat16):
        x = x32.to(dt).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = OF.dw_conv(x, w7, m75, m53, ks, True)
        g = torch.randn(y.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1)).to(y.dtype).contiguous(memory_format=torch.channels_last)
        w7.grad = None
        y.backward(g)
        res[dt] = (y.detach().float(), x.grad.detach().float(), w7.grad.detach().float().clone())
    a, b = res[torch.float32], res[torch.bfloat16]
    print('dw k%d  y cos %.6f  dx cos %.6f  dw cos %.6f' % (ks, cos(a[0], b[0]), cos(a[1], b[1]), cos(a[2], b[2])), flush=True)
