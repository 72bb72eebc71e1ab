"""Host-side helpers of the elastic-MBConv SR path.

Every function here is integer / bookkeeping logic whose results must be
bit-identical to the reference, or thin torch glue.  Nothing
in this file computes on activations; that is the job of the CUDA library
(`backend.py`).

Reference sites restated here:
  make_divisible          ofa/imagenet_codebase/utils/pytorch_modules.py:12-29
  int2list                ofa/imagenet_codebase/utils/__init__.py:92-98
  sub_filter_start_end    ofa/imagenet_codebase/utils/__init__.py:84-89
  get_same_padding        ofa/utils.py:211-219
  build_activation        ofa/utils.py:242-314
  pixel_unshuffle         ofa/utils.py:383-410
  MyModule / MyNetwork    ofa/utils.py:84-185
  psnr / tensor2img_np / rgb2y   ofa/utils.py:27-34, sr_run_manager.py:567-597
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = [
    'make_divisible', 'int2list', 'sub_filter_start_end', 'get_same_padding', 'build_activation',
    'PixelUnshuffle', 'pixel_unshuffle', 'Hswish', 'Hsigmoid', 'SEModule', 'MyModule', 'MyNetwork', 'psnr',
    'tensor2img_np', 'rgb2y', 'get_net_device', 'AverageMeter',
]


def make_divisible(v, divisor, min_val=None):
    """Round `v` to the nearest multiple of `divisor`, never dropping more than 10 %."""
    floor = divisor if min_val is None else min_val
    rounded = int(v + divisor / 2) // divisor * divisor
    out = floor if rounded < floor else rounded
    if out < 0.9 * v:
        out += divisor
    return out


def int2list(val, repeat_time=1):
    """Scalar -> list of `repeat_time` copies.  Lists / ndarrays are returned AS IS (same object):
    callers in the reference mutate the returned list and that aliasing is
    observable, so it is preserved."""
    if isinstance(val, (list, np.ndarray)):
        return val
    if isinstance(val, tuple):
        return list(val)
    return [val] * repeat_time


def sub_filter_start_end(kernel_size, sub_kernel_size):
    """Index range of the centred `sub_kernel_size` window inside a `kernel_size` filter."""
    lo = kernel_size // 2 - sub_kernel_size // 2
    hi = lo + sub_kernel_size
    assert hi - lo == sub_kernel_size and sub_kernel_size % 2 == 1
    return lo, hi


def get_same_padding(kernel_size):
    if isinstance(kernel_size, tuple):
        assert len(kernel_size) == 2, 'invalid kernel size: %s' % (kernel_size,)
        return get_same_padding(kernel_size[0]), get_same_padding(kernel_size[1])
    assert isinstance(kernel_size, int), 'kernel size should be either `int` or `tuple`'
    assert kernel_size % 2 > 0, 'kernel size should be odd number'
    return kernel_size // 2


class Hswish(nn.Module):
    def __init__(self, inplace=True):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        return x * F.relu6(x + 3., inplace=self.inplace) / 6.


class Hsigmoid(nn.Module):
    def __init__(self, inplace=True):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        return F.relu6(x + 3., inplace=self.inplace) / 6.


class SEModule(nn.Module):
    """Squeeze-and-excite (ofa/utils.py:354-375): same children (`fc.reduce`, `fc.relu`, `fc.expand`,
    `fc.h_sigmoid`) so checkpoints interchange; forward is the library's pool / sliced-linear / scale kernels."""
    REDUCTION = 4

    def __init__(self, channel):
        super().__init__()
        from collections import OrderedDict
        self.channel = channel
        self.reduction = SEModule.REDUCTION
        num_mid = make_divisible(self.channel // self.reduction, divisor=8)
        self.fc = nn.Sequential(OrderedDict([
            ('reduce', nn.Conv2d(self.channel, num_mid, 1, 1, 0, bias=True)),
            ('relu', nn.ReLU(inplace=True)),
            ('expand', nn.Conv2d(num_mid, self.channel, 1, 1, 0, bias=True)),
            ('h_sigmoid', Hsigmoid(inplace=True)),
        ]))

    def _se(self, x, num_mid):
        from . import functional as OF
        return OF.squeeze_excite(x, self.fc.reduce.weight, self.fc.reduce.bias, self.fc.expand.weight,
                                 self.fc.expand.bias, num_mid)

    def forward(self, x):
        return self._se(x, self.fc.reduce.weight.shape[0])


def pixel_unshuffle(input, downscale_factor):
    """out[n, c*r*r + y*r + x, h, w] = in[n, c, h*r + y, w*r + x]  (the reference builds this
    as a one-hot grouped conv; the result equals torch's pixel_unshuffle bit for bit)."""
    return F.pixel_unshuffle(input, downscale_factor)


class PixelUnshuffle(nn.Module):
    def __init__(self, downscale_factor):
        super().__init__()
        self.downscale_factor = downscale_factor

    def forward(self, input):
        return pixel_unshuffle(input, self.downscale_factor)


def build_activation(act_func, inplace=True, upscale_factor=2):
    """Same vocabulary as the reference.  'pixelshuffle' / 'pixelunshuffle' are "activations" there
    (a ConvLayer's third op); they stay nn.Modules here so state_dict / module trees agree, but
    ConvLayer.forward recognises them and fuses them into the conv kernel's store."""
    table = {
        'relu': lambda: nn.ReLU(inplace=inplace),
        'relu6': lambda: nn.ReLU6(inplace=inplace),
        'tanh': nn.Tanh,
        'sigmoid': nn.Sigmoid,
        'h_swish': lambda: Hswish(inplace=inplace),
        'h_sigmoid': lambda: Hsigmoid(inplace=inplace),
        'lrelu': lambda: nn.LeakyReLU(0.1, inplace=inplace),
        'pixelshuffle': lambda: nn.PixelShuffle(upscale_factor=2),
        'pixelunshuffle': lambda: PixelUnshuffle(downscale_factor=upscale_factor),
    }
    if act_func is None:
        return None
    if act_func not in table:
        raise ValueError('do not support: %s' % act_func)
    return table[act_func]()


def _invalidate_packed_weights():
    """`.data` writes do not move Tensor._version: tell the operator layer its derived 16-bit weight copies are void
    (they are re-derived per forward pass anyway; this is the explicit hook)."""
    from . import functional as OF
    OF.invalidate_packed_weights()


class MyModule(nn.Module):
    def forward(self, x):
        raise NotImplementedError

    @property
    def module_str(self):
        raise NotImplementedError

    @property
    def config(self):
        raise NotImplementedError

    @staticmethod
    def build_from_config(config):
        raise NotImplementedError


class MyNetwork(MyModule):
    """BN-parameter plumbing, initialisation and parameter-group selection used by the run manager."""

    def zero_last_gamma(self):
        raise NotImplementedError

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        _invalidate_packed_weights()
        return out

    def _bn_modules(self):
        return [m for m in self.modules() if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d))]

    def set_bn_param(self, momentum, eps):
        for m in self._bn_modules():
            m.momentum, m.eps = momentum, eps

    def get_bn_param(self):
        for m in self._bn_modules():
            return {'momentum': m.momentum, 'eps': m.eps}
        return None

    def init_model(self, model_init):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                kh, kw = m.kernel_size
                if model_init == 'he_fout':
                    fan = kh * kw * m.out_channels
                elif model_init == 'he_fin':
                    fan = kh * kw * m.in_channels
                else:
                    raise NotImplementedError
                m.weight.data.normal_(0, math.sqrt(2. / fan))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                bound = 1. / math.sqrt(m.weight.size(1))
                m.weight.data.uniform_(-bound, bound)
                if m.bias is not None:
                    m.bias.data.zero_()
        _invalidate_packed_weights()

    def get_parameters(self, keys=None, mode='include', exclude_set=None):
        skip = exclude_set or {}
        if mode not in ('include', 'exclude'):
            raise ValueError('do not support: %s' % mode)
        for name, param in self.named_parameters():
            if name in skip:
                continue
            if keys is not None:
                hit = any(k in name for k in keys)
                if hit != (mode == 'include'):
                    continue
            yield param

    def weight_parameters(self, exclude_set=None):
        return self.get_parameters(exclude_set=exclude_set)


def get_net_device(net):
    return next(net.parameters()).device


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


# ---------------------------------------------------------------------------------------------
# SR quality metric (reference definition: clamp [0,1] -> x255 -> round -> BT.601 Y -> PSNR)
# ---------------------------------------------------------------------------------------------

def psnr(img1, img2):
    assert img1.dtype == img2.dtype == np.uint8
    diff = img1.astype(np.float64) - img2.astype(np.float64)
    mse = np.mean(diff ** 2)
    if mse == 0:
        return float('inf')
    return 20 * math.log10(255.0 / math.sqrt(mse))


def tensor2img_np(tensor, out_type=np.uint8, min_max=(0, 1)):
    """[C,H,W] or [1,C,H,W] tensor in `min_max` -> HWC uint8 image (rounded)."""
    t = tensor.detach().float().cpu().clamp(*min_max)
    t = (t - min_max[0]) / (min_max[1] - min_max[0])
    if t.dim() == 4:
        assert t.size(0) == 1, 'grid view of a batch is not part of the hot path'
        t = t[0]
    arr = t.numpy()
    if t.dim() == 3:
        arr = np.transpose(arr, (1, 2, 0))
    elif t.dim() != 2:
        raise TypeError('Only support 4D, 3D and 2D tensor. But received dimension = %d' % t.dim())
    if out_type == np.uint8:
        arr = (arr * 255.0).round()
    return arr.astype(out_type)


def rgb2y(img):
    assert img.dtype == np.uint8
    y = (np.dot(img[..., :3], [65.481, 128.553, 24.966]) / 255.0 + 16.0).round()
    return y.astype(np.uint8)
