"""CPU ORACLE of the SR data preparation — TEST INFRASTRUCTURE ONLY.

What the reference does per training sample (ofa/imagenet_codebase/data_providers/div2k_setxx.py):
  :166-171  RandomCrop(image_size) -> RandomHorizontalFlip -> RandomRotation(degrees=(-90, 90)) on the PIL image
  :288-298  L2 = Scale(1/2, BICUBIC)(H), L4 = Scale(1/4, BICUBIC)(H); ToTensor() on H, L2, L4
  :355-380  Scale: target size (int(h * f), int(w * f)), `img.resize(size, Image.BICUBIC)`
The arithmetic lives in a third-party dependency that is NOT in /root/reference: Pillow (unpinned in
requirements.txt through torchvision; 12.2.0 in the build container) — `Image.resize` = libImaging/Resample.c
(two-pass separable resampling on uint8 with 22-bit fixed-point coefficients, antialiased: the filter support
is scaled by the down-scaling factor), `Image.rotate` = the affine nearest-neighbour path of libImaging/Geometry.c
(16.16 fixed point), and torchvision's ToTensor (uint8 HWC -> float32 CHW / 255).  This file restates those
published algorithms in numpy integer / float64 arithmetic.

Pinned: tests/golden/make_golden_prep.py runs Pillow itself in the build container on seeded images and stores
its outputs in tests/golden/reference_prep.npz; tests/test_oracle_golden.py checks this file against them
bit for bit.  Imported only by tests/ (never by the product package).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients are scaled by 2^22 for 8-bit images


def _bicubic(x, a=-0.5):
    """Resample.c bicubic_filter (Keys kernel, a = -0.5), double precision."""
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size, out_size, support=2.0):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole axis (box = [0, in_size)).
    Returns (ksize, bounds[out, 2] = (xmin, count), kk[out, ksize] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    sup = support * filterscale
    ksize = int(math.ceil(sup)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - sup + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + sup + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass_1d(img, out_size, axis):
    """One resampling pass along `axis` of a uint8 array: sum of pixel * coefficient in int32 from the rounding
    constant 2^21, arithmetic shift by 22, clip to [0, 255] (Resample.c ImagingResampleHorizontal/Vertical_8bpc)."""
    in_size = img.shape[axis]
    _, bounds, kk = resample_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def bicubic_resize_u8(img, out_h, out_w):
    """PIL `Image.resize((out_w, out_h), Image.BICUBIC)` on a uint8 HWC image: horizontal pass, then vertical
    pass on the uint8 intermediate (Resample.c ImagingResample)."""
    tmp = _pass_1d(img, out_w, 1) if out_w != img.shape[1] else img
    return _pass_1d(tmp, out_h, 0) if out_h != img.shape[0] else tmp


def scale_down(img, opt):
    """get_transform_L(opt) = Scale(1 / opt, BICUBIC) — div2k_setxx.py:355-380 (size = int(dim * scale))."""
    f = 1 / opt
    return bicubic_resize_u8(img, int(img.shape[0] * f), int(img.shape[1] * f))


def to_tensor(img):
    """torchvision ToTensor: uint8 HWC -> float32 CHW, `img.to(float32).div(255)`."""
    return (np.transpose(img, (2, 0, 1)).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def crop(img, i, j, h, w):
    """div2k_setxx.py:306-317 / torchvision RandomCrop's crop: rows i..i+h, columns j..j+w."""
    return img[i:i + h, j:j + w]


def hflip(img):
    """torchvision RandomHorizontalFlip -> PIL transpose(FLIP_LEFT_RIGHT)."""
    return img[:, ::-1]


def rotate_nearest(img, angle):
    """PIL `Image.rotate(angle)` with the defaults torchvision's RandomRotation passes (NEAREST, expand=False,
    centre = image centre, fill = 0): Image.py rotate() builds the inverse affine matrix in float64 with the
    sine / cosine rounded to 15 decimals, Geometry.c affine_fixed walks it in 16.16 fixed point."""
    h, w = img.shape[:2]
    angle = angle % 360.0
    if angle == 0:
        return img.copy()
    if angle == 180:
        return img[::-1, ::-1].copy()
    if angle in (90, 270) and h == w:
        return np.rot90(img, 1 if angle == 90 else 3).copy()
    cx, cy = w / 2.0, h / 2.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2] + cx
    m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5] + cy

    def fix(v):
        return int(math.floor(v * 65536.0 + 0.5))
    a0, a1, a3, a4 = fix(m[0]), fix(m[1]), fix(m[3]), fix(m[4])
    a2 = fix(m[2] + m[0] * 0.5 + m[1] * 0.5)
    a5 = fix(m[5] + m[3] * 0.5 + m[4] * 0.5)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.int64)
    xin = (a2 + a1 * ys + a0 * xs) >> 16
    yin = (a5 + a4 * ys + a3 * xs) >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.zeros_like(img)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def prepare_sample(img, i, j, size, flip, angle):
    """One sample of the training set (div2k_setxx.py:166-171, 288-298): crop -> flip -> rotate on the uint8 image,
    then the three tensors."""
    h = crop(img, i, j, size, size)
    if flip:
        h = hflip(h)
    h = rotate_nearest(np.ascontiguousarray(h), angle)
    return {'image': to_tensor(h), '2x_down_image': to_tensor(scale_down(h, 2)),
            '4x_down_image': to_tensor(scale_down(h, 4))}
