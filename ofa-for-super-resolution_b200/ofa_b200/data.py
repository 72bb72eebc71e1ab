"""SR data preparation on the device (training side).

The reference prepares every training sample in CPU data-loader workers with Pillow / torchvision
(ofa/imagenet_codebase/data_providers/div2k_setxx.py):
    :166-171  train transform  RandomCrop(image_size) -> RandomHorizontalFlip() -> RandomRotation((-90, 90))
    :288-298  __getitem__      H = transform(img); L2 = Scale(1/2, BICUBIC)(H); L4 = Scale(1/4, BICUBIC)(H);
                               {'image': ToTensor(H), '2x_down_image': ToTensor(L2), '4x_down_image': ToTensor(L4)}
    :355-380  Scale            size = (int(h * f), int(w * f)); img.resize(size[::-1], Image.BICUBIC)
Here a batch of uint8 RGB source images already resident in HBM goes through the same chain in three library calls
(augment, two resizes), bit-exact against Pillow (tests/golden/reference_prep.npz): the random parameters are drawn
on the host with the same torch RNG calls, in the same order, as the torchvision transforms draw them.
"""
import math

import torch

from . import backend as B


class _ResampleTable:
    """Pillow's (window, 22-bit coefficient) table for resizing one axis, built by the library's host function and
    uploaded once; owned by the caller (the library keeps no cache)."""

    def __init__(self, in_size, out_size, device):
        L = B.lib()
        self.ksize = int(L.ofa_resample_ksize(in_size, out_size))
        bounds = torch.empty((out_size, 2), dtype=torch.int32)
        kk = torch.empty((out_size, self.ksize), dtype=torch.int32)
        B.check(L.ofa_resample_build_table(in_size, out_size, bounds.data_ptr(), kk.data_ptr()))
        self.bounds_host, self.kk_host = bounds, kk
        self.bounds = bounds.to(device) if device is not None else None
        self.kk = kk.to(device) if device is not None else None


_tables = {}


def resample_table(in_size, out_size, device):
    key = (in_size, out_size, str(device))
    t = _tables.get(key)
    if t is None:
        t = _tables[key] = _ResampleTable(in_size, out_size, device)
    return t


def _check_u8(img):
    if not img.is_cuda:
        raise RuntimeError('libofa_sr_b200 has no CPU path: image batch is on %s' % img.device)
    assert img.dtype == torch.uint8 and img.dim() == 4 and img.shape[-1] == 3 and img.is_contiguous(), \
        'expected a contiguous uint8 [N, H, W, 3] batch'


def bicubic_resize(img, out_h, out_w, want_u8=False):
    """`PIL.Image.resize((out_w, out_h), Image.BICUBIC)` + ToTensor for a uint8 [N,H,W,3] device batch.
    Returns the fp32 [N,3,out_h,out_w] tensor (and the uint8 [N,out_h,out_w,3] image when want_u8)."""
    _check_u8(img)
    n, h, w, _ = img.shape
    th, tv = resample_table(w, out_w, img.device), resample_table(h, out_h, img.device)
    tmp = torch.empty((n, h, out_w, 3), dtype=torch.uint8, device=img.device)
    out = torch.empty((n, 3, out_h, out_w), dtype=torch.float32, device=img.device)
    out_u8 = torch.empty((n, out_h, out_w, 3), dtype=torch.uint8, device=img.device) if want_u8 else None
    B.check(B.lib().ofa_bicubic_resize_u8(img.data_ptr(), n, h, w, out_h, out_w, th.bounds.data_ptr(),
                                          th.kk.data_ptr(), th.ksize, tv.bounds.data_ptr(), tv.kk.data_ptr(),
                                          tv.ksize, tmp.data_ptr(), out_u8.data_ptr() if want_u8 else None,
                                          out.data_ptr(), B.stream_ptr(img.device)))
    return (out, out_u8) if want_u8 else out


def scale_down(img, opt):
    """get_transform_L(opt) + ToTensor — div2k_setxx.py:355-385."""
    assert opt in (2, 4, 8)
    f = 1 / opt
    return bicubic_resize(img, int(img.shape[1] * f), int(img.shape[2] * f))


def rotation_params(angle, size):
    """(mode, a0..a5) of `PIL.Image.rotate(angle)` (NEAREST, expand=False, centre = image centre) on a size x size
    patch: Image.py rotate() special-cases 0 / 180 / 90 / 270 degrees as transposes, otherwise builds the inverse
    matrix with sine and cosine rounded to 15 decimals, which Geometry.c walks in 16.16 fixed point."""
    angle = angle % 360.0
    if angle == 0:
        return (0, 0, 0, 0, 0, 0, 0)
    if angle == 180:
        return (1, 0, 0, 0, 0, 0, 0)
    if angle == 90:
        return (2, 0, 0, 0, 0, 0, 0)
    if angle == 270:
        return (3, 0, 0, 0, 0, 0, 0)
    c = size / 2.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * (-c) + m[1] * (-c) + m[2] + c
    m[5] = m[3] * (-c) + m[4] * (-c) + m[5] + c

    def fix(v):
        return int(math.floor(v * 65536.0 + 0.5))
    return (4, fix(m[0]), fix(m[1]), fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
            fix(m[3]), fix(m[4]), fix(m[5] + m[3] * 0.5 + m[4] * 0.5))


def sample_train_params(n, h, w, size, degrees=(-90.0, 90.0), flip_p=0.5):
    """Per-sample (i, j, flip, angle) drawn with the torch RNG calls torchvision's transforms make, in their order:
    RandomCrop.get_params (two randint), RandomHorizontalFlip (one rand), RandomRotation.get_params (one uniform_)."""
    out = []
    for _ in range(n):
        if h == size and w == size:
            i = j = 0
        else:
            i = int(torch.randint(0, h - size + 1, size=(1,)).item())
            j = int(torch.randint(0, w - size + 1, size=(1,)).item())
        flip = bool(torch.rand(1) < flip_p)
        angle = float(torch.empty(1).uniform_(float(degrees[0]), float(degrees[1])).item())
        out.append((i, j, flip, angle))
    return out


def augment(src, params, size, want_u8=True):
    """crop -> flip -> rotate (div2k_setxx.py:166-171) of one patch per source image.  src: uint8 [N,H,W,3] on the
    device; params: N tuples (i, j, flip, angle).  Returns (fp32 [N,3,size,size], uint8 [N,size,size,3])."""
    _check_u8(src)
    n, h, w, _ = src.shape
    assert len(params) == n
    rows = [(int(i), int(j), int(bool(flip))) + rotation_params(angle, size) for (i, j, flip, angle) in params]
    for (i, j, _f, *_r) in rows:
        if i < 0 or j < 0 or i + size > h or j + size > w:
            raise ValueError('crop (%d, %d) + %d leaves the %dx%d image' % (i, j, size, h, w))
    ptab = torch.tensor(rows, dtype=torch.int32).reshape(n, 10).to(src.device)
    out = torch.empty((n, 3, size, size), dtype=torch.float32, device=src.device)
    out_u8 = torch.empty((n, size, size, 3), dtype=torch.uint8, device=src.device) if want_u8 else None
    B.check(B.lib().ofa_sr_augment_u8(src.data_ptr(), h * w * 3, n, h, w, ptab.data_ptr(), size,
                                      out_u8.data_ptr() if want_u8 else None, out.data_ptr(),
                                      B.stream_ptr(src.device)))
    return out, out_u8


class SRTrainBatchPrep:
    """Device-side stand-in for the reference's training `Dataset.__getitem__` + collate: from a batch of uint8 source
    images to the dict of tensors `progressive_shrinking.train_one_epoch` consumes (keys as div2k_setxx.py:297)."""

    def __init__(self, image_size):
        self.image_size = image_size

    def __call__(self, src, params=None):
        n, h, w, _ = src.shape
        if params is None:
            params = sample_train_params(n, h, w, self.image_size)
        hr, hr_u8 = augment(src, params, self.image_size)
        return {'image': hr, '2x_down_image': scale_down(hr_u8, 2), '4x_down_image': scale_down(hr_u8, 4)}
