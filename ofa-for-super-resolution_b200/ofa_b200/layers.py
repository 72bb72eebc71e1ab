"""Static layers of the SR nets with the reference's module API (ofa/layers.py), backed by the
sm_100a kernels.

  ConvLayer              layers.py:120-196   conv -> BN -> {None | act | PixelShuffle | PixelUnshuffle}
  IdentityLayer          layers.py:310-332
  ZeroLayer              layers.py:419-444
  MBInvertedConvLayer    layers.py:447-526   the static twin DynamicMBConvLayer.get_active_subnet builds
  MobileInvertedResidualBlock   imagenet_codebase/networks/proxyless_nets.py:36-76

Module trees and state_dict keys are identical to the reference (`conv.weight`, `bn.*`); the
`forward`s do not walk the children one ATen call at a time but issue fused kernels.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from .utils import MyModule, build_activation, get_same_padding, SEModule
from . import functional as OF
from . import backend as B

__all__ = ['set_layer_from_config', 'ConvLayer', 'IdentityLayer', 'LinearLayer', 'ZeroLayer', 'MBInvertedConvLayer',
           'MobileInvertedResidualBlock']

_FUSABLE_ACTS = {None: B.ACT_NONE, 'relu6': B.ACT_RELU6, 'relu': B.ACT_RELU, 'h_swish': B.ACT_HSWISH}
_STORE_ACTS = {'pixelshuffle': B.STORE_PIXELSHUFFLE2, 'pixelunshuffle': B.STORE_PIXELUNSHUFFLE2}


def set_layer_from_config(layer_config):
    if layer_config is None:
        return None
    name2layer = {
        ConvLayer.__name__: ConvLayer,
        IdentityLayer.__name__: IdentityLayer,
        LinearLayer.__name__: LinearLayer,
        ZeroLayer.__name__: ZeroLayer,
        MBInvertedConvLayer.__name__: MBInvertedConvLayer,
    }
    layer_config = dict(layer_config)
    layer_name = layer_config.pop('name')
    return name2layer[layer_name].build_from_config(layer_config)


def _split_act(act_func):
    """reference 'activation' vocabulary -> (kernel activation code, store mode)"""
    if act_func in _STORE_ACTS:
        return B.ACT_NONE, _STORE_ACTS[act_func]
    if act_func in _FUSABLE_ACTS:
        return _FUSABLE_ACTS[act_func], B.STORE_PLAIN
    raise NotImplementedError('activation %r has no fused B200 epilogue (SR nets use relu6 / pixel(un)shuffle)' % act_func)


class ConvLayer(MyModule):
    """conv k x k (same padding, stride 1) -> BatchNorm -> act, `ops_order='weight_bn_act'`."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1, groups=1, bias=False,
                 has_shuffle=False, use_bn=True, act_func='relu', dropout_rate=0, ops_order='weight_bn_act'):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.dilation = dilation
        self.groups = groups
        self.bias = bias
        self.has_shuffle = has_shuffle
        self.use_bn = use_bn
        self.act_func = act_func
        self.dropout_rate = dropout_rate
        self.ops_order = ops_order
        if ops_order != 'weight_bn_act' or groups != 1 or bias or dropout_rate > 0 or stride != 1 or dilation != 1:
            raise NotImplementedError(
                'the B200 ConvLayer covers what the SR nets build: weight_bn_act, groups=1, no bias, '
                'no dropout, stride 1, dilation 1')
        padding = get_same_padding(self.kernel_size)
        # registration order = the reference's (conv, bn, act) so state_dict / module trees agree
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              padding=padding, dilation=dilation, groups=groups, bias=bias)
        if self.use_bn:
            self.bn = nn.BatchNorm2d(out_channels)
        act = build_activation(self.act_func, True)
        if act is not None:
            self.act = act
        self._act_code, self._store = _split_act(self.act_func)
        self._packed = OF.PackedWeightCache()
        # set by the owning network on its last layer: hand the caller an fp32 NCHW tensor
        self.out_dtype = None
        self.out_nchw = False

    @OF.scoped_forward()
    def forward(self, x, residual=None):
        """`residual` (optional) is added after BN/act/shuffle — the networks use it to fuse their
        in-place long-skip adds (`x += dec_big_skip`, ofa_mbs4.py:159)."""
        bn = self.bn if self.use_bn else None
        w = self.conv.weight
        if OF.inference_mode_active(self) and not OF.bn_hooked(bn):
            out_dtype = self.out_dtype
            if out_dtype is None and self.out_channels < 16 and self._store == B.STORE_PLAIN:
                # thin tensors (the 3-channel learned low-resolution image of X4) stay fp32: it costs
                # nothing and a bf16 rounding at a 3-channel bottleneck perturbs the image directly
                out_dtype = torch.float32
            return OF.conv_bn_act_infer(x, w, self.in_channels, self.out_channels, self.kernel_size, bn,
                                        self._act_code, self._store, residual, self._packed,
                                        out_dtype, self.out_nchw)
        if self._store == B.STORE_PLAIN:
            if bn is not None:      # conv -> BN -> act (+ residual) as one autograd node
                y = OF.conv_bn_act(x, w, self.in_channels, self.out_channels, self.kernel_size, bn, self._act_code,
                                   residual)
            else:
                assert self._act_code == B.ACT_NONE and residual is None
                y = OF.conv2d(x, w, self.in_channels, self.out_channels, self.kernel_size)
            if self.out_dtype is not None and y.dtype != torch.float32:
                y = y.to(torch.float32)       # mixed-precision training: the caller's loss sees fp32 (whatever image
                                              # format set_output_dtype chose for inference)
            return y
        if bn is not None:
            y = OF.conv_bn_act(x, w, self.in_channels, self.out_channels, self.kernel_size, bn, B.ACT_NONE, None)
        else:
            y = OF.conv2d(x, w, self.in_channels, self.out_channels, self.kernel_size)
        y = OF.pixel_shuffle2(y) if self._store == B.STORE_PIXELSHUFFLE2 else OF.pixel_unshuffle2(y)
        if residual is not None:
            y = y + residual
        return y

    @property
    def module_str(self):
        k = self.kernel_size if isinstance(self.kernel_size, tuple) else (self.kernel_size, self.kernel_size)
        return '%dx%d_Conv_O%d' % (k[0], k[1], self.out_channels)

    @property
    def config(self):
        return {
            'name': ConvLayer.__name__,
            'kernel_size': self.kernel_size, 'stride': self.stride, 'dilation': self.dilation,
            'groups': self.groups, 'bias': self.bias, 'has_shuffle': self.has_shuffle,
            'in_channels': self.in_channels, 'out_channels': self.out_channels, 'use_bn': self.use_bn,
            'act_func': self.act_func, 'dropout_rate': self.dropout_rate, 'ops_order': self.ops_order,
        }

    @staticmethod
    def build_from_config(config):
        return ConvLayer(**config)


class IdentityLayer(MyModule):
    def __init__(self, in_channels, out_channels, use_bn=False, act_func=None, dropout_rate=0,
                 ops_order='weight_bn_act'):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.use_bn = use_bn
        self.act_func = act_func
        self.dropout_rate = dropout_rate
        self.ops_order = ops_order
        if use_bn or act_func is not None or dropout_rate > 0:
            raise NotImplementedError('IdentityLayer with BN / activation / dropout is not used by the SR nets')

    def forward(self, x):
        return x

    @property
    def module_str(self):
        return 'Identity'

    @property
    def config(self):
        return {
            'name': IdentityLayer.__name__,
            'in_channels': self.in_channels, 'out_channels': self.out_channels, 'use_bn': self.use_bn,
            'act_func': self.act_func, 'dropout_rate': self.dropout_rate, 'ops_order': self.ops_order,
        }

    @staticmethod
    def build_from_config(config):
        return IdentityLayer(**config)


class LinearLayer(MyModule):
    """Static fully connected layer (ofa/layers.py:335-418) in the only configuration the elastic nets build —
    [dropout ->] linear, no BatchNorm1d, no activation (DynamicLinearLayer.get_active_subnet, classifier heads) —
    on the library's sliced-linear kernel."""

    def __init__(self, in_features, out_features, bias=True, use_bn=False, act_func=None, dropout_rate=0,
                 ops_order='weight_bn_act'):
        super().__init__()
        if use_bn or act_func is not None:
            raise NotImplementedError('LinearLayer with BatchNorm1d / activation is not used by the elastic nets')
        self.in_features = in_features
        self.out_features = out_features
        self.bias = bias
        self.use_bn = use_bn
        self.act_func = act_func
        self.dropout_rate = dropout_rate
        self.ops_order = ops_order
        if self.dropout_rate > 0:
            self.dropout = nn.Dropout(self.dropout_rate, inplace=True)
        self.linear = nn.Linear(self.in_features, self.out_features, self.bias)

    def forward(self, x):
        if self.dropout_rate > 0:
            x = self.dropout(x)
        return OF.linear(x, self.linear.weight, self.linear.bias if self.bias else None, self.out_features)

    @property
    def module_str(self):
        return '%dx%d_Linear' % (self.in_features, self.out_features)

    @property
    def config(self):
        return {
            'name': LinearLayer.__name__, 'in_features': self.in_features, 'out_features': self.out_features,
            'bias': self.bias, 'use_bn': self.use_bn, 'act_func': self.act_func, 'dropout_rate': self.dropout_rate,
            'ops_order': self.ops_order,
        }

    @staticmethod
    def build_from_config(config):
        return LinearLayer(**config)


class ZeroLayer(MyModule):
    def __init__(self, stride):
        super().__init__()
        self.stride = stride

    def forward(self, x):
        raise ValueError

    @property
    def module_str(self):
        return 'Zero'

    @property
    def config(self):
        return {'name': ZeroLayer.__name__, 'stride': self.stride}

    @staticmethod
    def build_from_config(config):
        return ZeroLayer(**config)


class MBInvertedConvLayer(MyModule):
    """Static MBConv (expand 1x1 -> depthwise k x k -> project 1x1, each + BN, ReLU6 after the first
    two).  Built by DynamicMBConvLayer.get_active_subnet; a second statement of what the active
    sub-network computes (SURVEY §8c 'secondary oracle')."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, expand_ratio=6, mid_channels=None,
                 act_func='relu6', use_se=False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.expand_ratio = expand_ratio
        self.mid_channels = mid_channels
        self.act_func = act_func
        self.use_se = use_se
        feature_dim = round(self.in_channels * self.expand_ratio) if self.mid_channels is None else self.mid_channels
        self._feature_dim = feature_dim
        if self.expand_ratio == 1:
            self.inverted_bottleneck = None
        else:
            self.inverted_bottleneck = nn.Sequential(OrderedDict([
                ('conv', nn.Conv2d(self.in_channels, feature_dim, 1, 1, 0, bias=False)),
                ('bn', nn.BatchNorm2d(feature_dim)),
                ('act', build_activation(self.act_func, inplace=True)),
            ]))
        pad = get_same_padding(self.kernel_size)
        depth_conv_modules = [
            ('conv', nn.Conv2d(feature_dim, feature_dim, kernel_size, stride, pad, groups=feature_dim, bias=False)),
            ('bn', nn.BatchNorm2d(feature_dim)),
            ('act', build_activation(self.act_func, inplace=True)),
        ]
        if self.use_se:
            depth_conv_modules.append(('se', SEModule(feature_dim)))
        self.depth_conv = nn.Sequential(OrderedDict(depth_conv_modules))
        self.point_linear = nn.Sequential(OrderedDict([
            ('conv', nn.Conv2d(feature_dim, out_channels, 1, 1, 0, bias=False)),
            ('bn', nn.BatchNorm2d(out_channels)),
        ]))
        self._act_code, _ = _split_act(self.act_func)
        self._packed_exp = OF.PackedWeightCache()
        self._packed_proj = OF.PackedWeightCache()

    @OF.scoped_forward()
    def forward(self, x, residual=None):
        act, mid = self._act_code, self._feature_dim
        infer = OF.inference_mode_active(self) and not OF.bn_hooked(
            self.inverted_bottleneck.bn if self.inverted_bottleneck is not None else None, self.depth_conv.bn,
            self.point_linear.bn)
        dw = self.depth_conv.conv.weight
        if self.use_se or self.stride != 1:
            # MobileNetV3 flavour: unfused kernels, same in train and eval mode
            if self.inverted_bottleneck is not None:
                x = OF.conv2d(x, self.inverted_bottleneck.conv.weight, self.in_channels, mid, 1)
                x = OF.bn_act(x, self.inverted_bottleneck.bn, mid, act)
            x = OF.dw_conv(x, dw, None, None, self.kernel_size, False, self.stride)
            x = OF.bn_act(x, self.depth_conv.bn, mid, act)
            if self.use_se:
                x = self.depth_conv.se(x)
            x = OF.conv2d(x, self.point_linear.conv.weight, mid, self.out_channels, 1)
            return OF.bn_act(x, self.point_linear.bn, self.out_channels, B.ACT_NONE, residual)
        if infer:
            if self.inverted_bottleneck is not None:
                x = OF.conv_bn_act_infer(x, self.inverted_bottleneck.conv.weight, self.in_channels, mid, 1,
                                         self.inverted_bottleneck.bn, act, cache=self._packed_exp)
            x = OF.dw_bn_act_infer(x, dw, None, None, self.kernel_size, False, self.depth_conv.bn, act)
            return OF.conv_bn_act_infer(x, self.point_linear.conv.weight, mid, self.out_channels, 1,
                                        self.point_linear.bn, B.ACT_NONE, residual=residual, cache=self._packed_proj)
        if self.inverted_bottleneck is not None:
            x = OF.conv2d(x, self.inverted_bottleneck.conv.weight, self.in_channels, mid, 1)
            x = OF.bn_act(x, self.inverted_bottleneck.bn, mid, act)
        x = OF.dw_conv(x, dw, None, None, self.kernel_size, False)
        x = OF.bn_act(x, self.depth_conv.bn, mid, act)
        x = OF.conv2d(x, self.point_linear.conv.weight, mid, self.out_channels, 1)
        return OF.bn_act(x, self.point_linear.bn, self.out_channels, B.ACT_NONE, residual)

    @property
    def module_str(self):
        expand_ratio = self.expand_ratio if self.mid_channels is None else self.mid_channels // self.in_channels
        s = '%dx%d_MBConv%d_%s' % (self.kernel_size, self.kernel_size, expand_ratio, self.act_func.upper())
        if self.use_se:
            s = 'SE_' + s
        return s + '_O%d' % self.out_channels

    @property
    def config(self):
        return {
            'name': MBInvertedConvLayer.__name__,
            'in_channels': self.in_channels, 'out_channels': self.out_channels,
            'kernel_size': self.kernel_size, 'stride': self.stride, 'expand_ratio': self.expand_ratio,
            'mid_channels': self.mid_channels, 'act_func': self.act_func, 'use_se': self.use_se,
        }

    @staticmethod
    def build_from_config(config):
        return MBInvertedConvLayer(**config)


class MobileInvertedResidualBlock(MyModule):
    """res = mobile_inverted_conv(x) + shortcut(x)  (proxyless_nets.py:44-51); with an IdentityLayer
    shortcut the add is fused into the project conv's epilogue."""

    def __init__(self, mobile_inverted_conv, shortcut):
        super().__init__()
        self.mobile_inverted_conv = mobile_inverted_conv
        self.shortcut = shortcut

    @OF.scoped_forward()
    def forward(self, x):
        if self.mobile_inverted_conv is None or isinstance(self.mobile_inverted_conv, ZeroLayer):
            return x
        if self.shortcut is None or isinstance(self.shortcut, ZeroLayer):
            return self.mobile_inverted_conv(x)
        return self.mobile_inverted_conv(x, residual=self.shortcut(x))

    @property
    def module_str(self):
        return '(%s, %s)' % (
            self.mobile_inverted_conv.module_str if self.mobile_inverted_conv is not None else None,
            self.shortcut.module_str if self.shortcut is not None else None,
        )

    @property
    def config(self):
        return {
            'name': MobileInvertedResidualBlock.__name__,
            'mobile_inverted_conv': self.mobile_inverted_conv.config if self.mobile_inverted_conv is not None else None,
            'shortcut': self.shortcut.config if self.shortcut is not None else None,
        }

    @staticmethod
    def build_from_config(config):
        mobile_inverted_conv = set_layer_from_config(config['mobile_inverted_conv'])
        shortcut = set_layer_from_config(config['shortcut'])
        return MobileInvertedResidualBlock(mobile_inverted_conv, shortcut)
