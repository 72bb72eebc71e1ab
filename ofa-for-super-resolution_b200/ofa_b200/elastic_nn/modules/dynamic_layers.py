"""Elastic MBConv block with the reference's API (ofa/elastic_nn/modules/dynamic_layers.py:14-199).

Children keep the reference names and nesting — `inverted_bottleneck.{conv,bn,act}`,
`depth_conv.{conv,bn,act}`, `point_linear.{conv,bn}` — so checkpoints interchange; `forward` issues
fused kernels instead of walking the Sequentials:
  inference : ONE library call for the block (expand -> BN -> ReLU6 -> elastic depthwise -> BN ->
              ReLU6 -> project -> BN [+ identity residual]) on NHWC bf16, or three fused
              conv+BN+act kernels on the exact fp32 path;
  training  : conv / BN(+act) autograd Functions whose forward and backward are library kernels.
"""
import copy
from collections import OrderedDict

import torch
import torch.nn as nn

from ...layers import MBInvertedConvLayer, ConvLayer, LinearLayer, _split_act
from ...utils import MyModule, int2list, get_net_device, build_activation, make_divisible, SEModule
from ..utils import adjust_bn_according_to_idx, copy_bn
from .dynamic_op import DynamicSeparableConv2d, DynamicPointConv2d, DynamicBatchNorm2d, DynamicLinear, DynamicSE
from ... import functional as OF
from ... import backend as B

__all__ = ['DynamicMBConvLayer', 'DynamicConvLayer', 'DynamicLinearLayer']


def _fill_from_slice(dst, src, *extent):
    """dst <- src[:extent[0], :extent[1], ...] (the active prefix of a full-size parameter)."""
    dst.data.copy_(src.data[tuple(slice(0, e) for e in extent)])


def _permute(param, dim, order):
    """Reorder a parameter along `dim` in place (the tensor object is replaced, the Parameter stays)."""
    param.data = torch.index_select(param.data, dim, order)


def _has_hooks(*mods):
    """True when a sub-module carries forward / backward hooks (then it is called as a module, not bypassed)."""
    for m in mods:
        if m is not None and (m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or
                              getattr(m, '_backward_pre_hooks', None)):
            return True
    return False


class DynamicMBConvLayer(MyModule):

    def __init__(self, in_channel_list, out_channel_list, kernel_size_list=3, expand_ratio_list=6, stride=1,
                 act_func='relu6', use_se=False):
        super().__init__()
        self.in_channel_list = in_channel_list
        self.out_channel_list = out_channel_list
        self.kernel_size_list = int2list(kernel_size_list, 1)
        self.expand_ratio_list = int2list(expand_ratio_list, 1)
        self.stride = stride
        self.act_func = act_func
        self.use_se = use_se

        max_middle_channel = round(max(self.in_channel_list) * max(self.expand_ratio_list))
        has_act = build_activation(self.act_func, inplace=True) is not None

        def seq(conv, width):
            mods = [('conv', conv), ('bn', DynamicBatchNorm2d(width))]
            if has_act:
                mods.append(('act', build_activation(self.act_func, inplace=True)))
            return nn.Sequential(OrderedDict(mods))

        if max(self.expand_ratio_list) == 1:
            self.inverted_bottleneck = None
        else:
            self.inverted_bottleneck = seq(DynamicPointConv2d(max(self.in_channel_list), max_middle_channel),
                                           max_middle_channel)
        self.depth_conv = seq(DynamicSeparableConv2d(max_middle_channel, self.kernel_size_list, self.stride),
                              max_middle_channel)
        if self.use_se:
            self.depth_conv.add_module('se', DynamicSE(max_middle_channel))
        self.point_linear = nn.Sequential(OrderedDict([
            ('conv', DynamicPointConv2d(max_middle_channel, max(self.out_channel_list))),
            ('bn', DynamicBatchNorm2d(max(self.out_channel_list))),
        ]))

        self.active_kernel_size = max(self.kernel_size_list)
        self.active_expand_ratio = max(self.expand_ratio_list)
        self.active_out_channel = max(self.out_channel_list)
        self._act_code, _ = _split_act(self.act_func)

    # ---------------------------------------------------------------------------------------------
    @OF.scoped_forward()
    def forward(self, x, residual=None):
        in_channel = x.size(1)
        if self.inverted_bottleneck is not None:
            self.inverted_bottleneck.conv.active_out_channel = \
                make_divisible(round(in_channel * self.active_expand_ratio), 8)
        self.depth_conv.conv.active_kernel_size = self.active_kernel_size
        self.point_linear.conv.active_out_channel = self.active_out_channel

        act = self._act_code
        ks = self.active_kernel_size
        cout = self.active_out_channel
        dwm = self.depth_conv.conv
        dwm._check_supported(ks)
        m75, m53 = dwm._matrices()
        transform_on = DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE is not None
        bn_dw = self.depth_conv.bn.bn
        bn_pl = self.point_linear.bn.bn
        w_pl = self.point_linear.conv.conv.weight
        bn_ex = self.inverted_bottleneck.bn.bn if self.inverted_bottleneck is not None else None
        hooked = DynamicBatchNorm2d.SET_RUNNING_STATISTICS or OF.bn_hooked(bn_ex, bn_dw, bn_pl)

        if OF.inference_mode_active(self) and not hooked and not self.use_se and self.stride == 1:
            if self.inverted_bottleneck is not None:
                mid = self.inverted_bottleneck.conv.active_out_channel
                exp = self.inverted_bottleneck.conv
                bn_exp = self.inverted_bottleneck.bn.bn
                fused_ok = (x.dtype in (torch.bfloat16, torch.float16) and x.is_contiguous(memory_format=torch.channels_last)
                            and in_channel % 64 == 0 and mid % 64 == 0 and cout % 64 == 0
                            and (residual is None or residual is x) and OF._state['impl'] != B.IMPL_SIMT)
                if fused_ok:
                    return OF.mbconv_infer(x, exp.conv.weight, dwm.conv.weight, m75, m53, w_pl, in_channel, mid,
                                           cout, ks, transform_on, act, bn_exp, bn_dw, bn_pl, residual is not None,
                                           pack_cache=self.__dict__.setdefault('_planar_pack', {}))
                h = OF.conv_bn_act_infer(x, exp.conv.weight, in_channel, mid, 1, bn_exp, act, cache=exp._packed)
            else:
                mid, h = in_channel, x
            h = OF.dw_bn_act_infer(h, dwm.conv.weight, m75, m53, ks, transform_on, bn_dw, act)
            return OF.conv_bn_act_infer(h, w_pl, mid, cout, 1, bn_pl, B.ACT_NONE, residual=residual,
                                        cache=self.point_linear.conv._packed)

        # autograd path (training, or eval with grad enabled)
        h = x
        mid = in_channel
        if self.inverted_bottleneck is not None:
            mid = self.inverted_bottleneck.conv.active_out_channel
        if self.stride == 1 and not _has_hooks(self.inverted_bottleneck.conv if self.inverted_bottleneck is not None
                                               else None, dwm, self.point_linear.conv):
            if (self.inverted_bottleneck is not None and not self.use_se and OF._state['block_train']
                    and OF.mbconv_train_supported(x, in_channel, mid, cout, residual, (bn_ex, bn_dw, bn_pl))):
                # the whole block as one autograd node / one library call each way (the eager step is host-bound)
                return OF.mbconv_train(x, self.inverted_bottleneck.conv.conv.weight, dwm.conv.weight, m75, m53, w_pl, mid,
                                       cout, ks, transform_on, act, bn_ex, bn_dw, bn_pl, residual is not None)
            # one autograd node per conv + BN + act layer (same library calls; the eager step is launch-rate-bound)
            if self.inverted_bottleneck is not None:
                h = OF.conv_bn_act(h, self.inverted_bottleneck.conv.conv.weight, in_channel, mid, 1, bn_ex, act)
            h = OF.dw_bn_act(h, dwm.conv.weight, m75, m53, ks, transform_on, bn_dw, act)
            if self.use_se:
                h = self.depth_conv.se(h)
            return OF.conv_bn_act(h, w_pl, mid, cout, 1, bn_pl, B.ACT_NONE, residual)
        if self.inverted_bottleneck is not None:
            h = self.inverted_bottleneck.conv(h)
            h = DynamicBatchNorm2d.bn_forward(h, self.inverted_bottleneck.bn.bn, mid, act)
        h = dwm(h)
        h = DynamicBatchNorm2d.bn_forward(h, bn_dw, mid, act)
        if self.use_se:
            h = self.depth_conv.se(h)
        h = self.point_linear.conv(h)
        return DynamicBatchNorm2d.bn_forward(h, bn_pl, cout, B.ACT_NONE, residual)

    @property
    def module_str(self):
        core = '(O%d, E%.1f, K%d)' % (self.active_out_channel, self.active_expand_ratio, self.active_kernel_size)
        return 'SE' + core if self.use_se else core

    @property
    def config(self):
        return {
            'name': DynamicMBConvLayer.__name__,
            'in_channel_list': self.in_channel_list, 'out_channel_list': self.out_channel_list,
            'kernel_size_list': self.kernel_size_list, 'expand_ratio_list': self.expand_ratio_list,
            'stride': self.stride, 'act_func': self.act_func, 'use_se': self.use_se,
        }

    @staticmethod
    def build_from_config(config):
        return DynamicMBConvLayer(**config)

    # ---------------------------------------------------------------------------------------------
    def get_active_subnet(self, in_channel, preserve_weight=True):
        """Static MBInvertedConvLayer holding exactly the active weights (dynamic_layers.py:112-154)."""
        middle_channel = make_divisible(round(in_channel * self.active_expand_ratio), 8)
        sub_layer = MBInvertedConvLayer(
            in_channel, self.active_out_channel, self.active_kernel_size, self.stride, self.active_expand_ratio,
            act_func=self.act_func, mid_channels=middle_channel, use_se=self.use_se,
        ).to(get_net_device(self))
        if not preserve_weight:
            return sub_layer
        mid, cout = middle_channel, self.active_out_channel
        if sub_layer.inverted_bottleneck is not None:
            _fill_from_slice(sub_layer.inverted_bottleneck.conv.weight, self.inverted_bottleneck.conv.conv.weight,
                             mid, in_channel)
            copy_bn(sub_layer.inverted_bottleneck.bn, self.inverted_bottleneck.bn.bn)
        # the depthwise filter of the sub-layer is the TRANSFORMED active filter, not a slice
        sub_layer.depth_conv.conv.weight.data.copy_(self.depth_conv.conv.get_active_filter(mid, self.active_kernel_size).data)
        copy_bn(sub_layer.depth_conv.bn, self.depth_conv.bn.bn)
        if self.use_se:
            se_mid = make_divisible(mid // SEModule.REDUCTION, divisor=8)
            full, part = self.depth_conv.se.fc, sub_layer.depth_conv.se.fc
            _fill_from_slice(part.reduce.weight, full.reduce.weight, se_mid, mid)
            _fill_from_slice(part.reduce.bias, full.reduce.bias, se_mid)
            _fill_from_slice(part.expand.weight, full.expand.weight, mid, se_mid)
            _fill_from_slice(part.expand.bias, full.expand.bias, mid)
        _fill_from_slice(sub_layer.point_linear.conv.weight, self.point_linear.conv.conv.weight, cout, mid)
        copy_bn(sub_layer.point_linear.bn, self.point_linear.bn.bn)
        return sub_layer

    def re_organize_middle_weights(self, expand_ratio_stage=0):
        """Sort the middle channels by the L1 norm of the project weights, descending
        (dynamic_layers.py:156-199) so every narrower expand ratio keeps the most important ones."""
        OF.invalidate_packed_weights()
        importance = torch.sum(torch.abs(self.point_linear.conv.conv.weight.data), dim=(0, 2, 3))
        if expand_ratio_stage > 0:
            sorted_expand_list = sorted(copy.deepcopy(self.expand_ratio_list), reverse=True)
            target_width = round(max(self.in_channel_list) * sorted_expand_list[expand_ratio_stage])
            importance[target_width:] = torch.arange(0, target_width - importance.size(0), -1)
        _, sorted_idx = torch.sort(importance, dim=0, descending=True)
        # every tensor indexed by the middle channel follows the new order
        _permute(self.point_linear.conv.conv.weight, 1, sorted_idx)
        adjust_bn_according_to_idx(self.depth_conv.bn.bn, sorted_idx)
        _permute(self.depth_conv.conv.conv.weight, 0, sorted_idx)
        if self.use_se:
            fc = self.depth_conv.se.fc
            _permute(fc.expand.weight, 0, sorted_idx)
            _permute(fc.expand.bias, 0, sorted_idx)
            _permute(fc.reduce.weight, 1, sorted_idx)
            # ... and the squeeze channels are ranked by their own importance (dynamic_layers.py:175-189)
            se_rank = torch.sort(torch.sum(torch.abs(fc.expand.weight.data), dim=(0, 2, 3)), dim=0, descending=True)[1]
            _permute(fc.expand.weight, 1, se_rank)
            _permute(fc.reduce.weight, 0, se_rank)
            _permute(fc.reduce.bias, 0, se_rank)
        if self.inverted_bottleneck is None:
            return sorted_idx
        adjust_bn_according_to_idx(self.inverted_bottleneck.bn.bn, sorted_idx)
        _permute(self.inverted_bottleneck.conv.conv.weight, 0, sorted_idx)
        return None


class DynamicConvLayer(MyModule):
    """Elastic-width conv -> BN -> act (dynamic_layers.py:202-270); not instantiated by the SR nets but
    part of the ofa/elastic_nn surface the modules above share."""

    def __init__(self, in_channel_list, out_channel_list, kernel_size=3, stride=1, dilation=1, use_bn=True,
                 act_func='relu6'):
        super().__init__()
        self.in_channel_list = in_channel_list
        self.out_channel_list = out_channel_list
        self.kernel_size = kernel_size
        self.stride = stride
        self.dilation = dilation
        self.use_bn = use_bn
        self.act_func = act_func
        self.conv = DynamicPointConv2d(
            max_in_channels=max(self.in_channel_list), max_out_channels=max(self.out_channel_list),
            kernel_size=self.kernel_size, stride=self.stride, dilation=self.dilation,
        )
        if self.use_bn:
            self.bn = DynamicBatchNorm2d(max(self.out_channel_list))
        self.act = build_activation(self.act_func, inplace=True)
        self.active_out_channel = max(self.out_channel_list)
        self._act_code, _ = _split_act(self.act_func)

    @OF.scoped_forward()
    def forward(self, x):
        self.conv.active_out_channel = self.active_out_channel
        cin, cout = x.size(1), self.active_out_channel
        bn = self.bn.bn if self.use_bn else None
        if OF.inference_mode_active(self) and self.dilation == 1:
            s = self.stride
            if s != 1 and self.kernel_size == 1:
                x, s = x[:, :, ::s, ::s], 1
            y = OF.conv_bn_act_infer(x, self.conv.conv.weight, cin, cout, self.kernel_size, bn, self._act_code,
                                     cache=self.conv._packed)
            return y if s == 1 else y[:, :, ::s, ::s]     # BN and the activation are pointwise: subsample afterwards
        y = self.conv(x)
        if bn is not None:
            return DynamicBatchNorm2d.bn_forward(y, bn, cout, self._act_code)
        assert self._act_code == B.ACT_NONE
        return y

    @property
    def module_str(self):
        return 'DyConv(O%d, K%d, S%d)' % (self.active_out_channel, self.kernel_size, self.stride)

    @property
    def config(self):
        return {
            'name': DynamicConvLayer.__name__,
            'in_channel_list': self.in_channel_list, 'out_channel_list': self.out_channel_list,
            'kernel_size': self.kernel_size, 'stride': self.stride, 'dilation': self.dilation,
            'use_bn': self.use_bn, 'act_func': self.act_func,
        }

    @staticmethod
    def build_from_config(config):
        return DynamicConvLayer(**config)

    def get_active_subnet(self, in_channel, preserve_weight=True):
        sub_layer = ConvLayer(in_channel, self.active_out_channel, self.kernel_size, self.stride, self.dilation,
                              use_bn=self.use_bn, act_func=self.act_func).to(get_net_device(self))
        if not preserve_weight:
            return sub_layer
        sub_layer.conv.weight.data.copy_(self.conv.conv.weight.data[:self.active_out_channel, :in_channel, :, :])
        if self.use_bn:
            copy_bn(sub_layer.bn, self.bn.bn)
        return sub_layer


from .dynamic_linear import DynamicLinearLayer  # noqa: E402,F401  (re-exported: the reference defines it in this module)
