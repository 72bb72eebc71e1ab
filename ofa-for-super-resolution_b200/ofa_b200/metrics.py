"""Evaluation metric of the SR run manager on the device.

`psnr_y(output, target)` returns what the reference's validate loop computes per batch —
`psnr(rgb2y(tensor2img_np(output)), rgb2y(tensor2img_np(images)))` (sr_run_manager.py:364,496,567-597;
ofa/utils.py:27-34) — without moving the images to the host: the kernel produces the exact integer sum of
squared luma differences per image, and only N int64 values cross PCIe.
"""
import math
from ctypes import byref

import torch

from . import backend as B


def psnr_y_sse(a, b):
    """Per-image sum over pixels of (Y_a - Y_b)^2, Y = BT.601 luma of the uint8-rounded image.  int64 [N] (device)."""
    assert a.shape == b.shape and a.dim() == 4 and a.shape[1] == 3, 'expected two [N,3,H,W] batches'
    sse = torch.zeros(a.shape[0], dtype=torch.int64, device=a.device)
    if a.numel():
        ta, tb = B.t4(a), B.t4(b)
        B.check(B.lib().ofa_psnr_y_sse(byref(ta), byref(tb), sse.data_ptr(), B.stream_ptr(a.device)))
    return sse


def psnr_y(a, b):
    """The reference metric, including its batch quirk: for N > 1 tensor2img_np tiles the batch with
    torchvision.utils.make_grid(nrow=int(sqrt(N)), padding=2) and the mean runs over the whole grid."""
    n, _, h, w = a.shape
    total = int(psnr_y_sse(a, b).sum().item())
    if n == 1:
        count = h * w
    else:
        xmaps = min(int(math.sqrt(n)), n)
        ymaps = int(math.ceil(float(n) / xmaps))
        count = ((h + 2) * ymaps + 2) * ((w + 2) * xmaps + 2)
    mse = total / count
    if mse == 0:
        return float('inf')
    return 20 * math.log10(255.0 / math.sqrt(mse))
