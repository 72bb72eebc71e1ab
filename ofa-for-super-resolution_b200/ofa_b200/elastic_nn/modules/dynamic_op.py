"""Elastic primitive ops with the reference's module API, backed by the sm_100a kernels.

Mirrors ofa/elastic_nn/modules/dynamic_op.py of the reference (class names, constructor signatures,
attribute and parameter names, class-level switches) so state_dicts interchange and callers are
unchanged:
  DynamicSeparableConv2d   dynamic_op.py:14-84    kernel-size-elastic depthwise conv
  DynamicPointConv2d       dynamic_op.py:87-112   channel-sliced dense conv (1x1 in the SR nets)
  DynamicBatchNorm2d       dynamic_op.py:139-172  BatchNorm over the active channel prefix
The parameters stay ordinary fp32 nn.Parameters at full supernet size; kernels address the active
slice in place (no `.contiguous()` copies).
"""
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from ...utils import get_same_padding, sub_filter_start_end, make_divisible, SEModule
from ... import functional as OF
from ... import backend as B

__all__ = ['DynamicSeparableConv2d', 'DynamicPointConv2d', 'DynamicBatchNorm2d', 'DynamicLinear', 'DynamicSE']


class DynamicSeparableConv2d(nn.Module):
    KERNEL_TRANSFORM_MODE = None  # None or 1 — read at construction AND at forward, like the reference

    def __init__(self, max_in_channels, kernel_size_list, stride=1, dilation=1):
        super().__init__()
        self.max_in_channels = max_in_channels
        self.kernel_size_list = kernel_size_list
        self.stride = stride
        self.dilation = dilation

        self.conv = nn.Conv2d(
            self.max_in_channels, self.max_in_channels, max(self.kernel_size_list), self.stride,
            groups=self.max_in_channels, bias=False,
        )
        self._ks_set = sorted(set(self.kernel_size_list))
        if self.KERNEL_TRANSFORM_MODE is not None:
            for small, large in zip(self._ks_set[:-1], self._ks_set[1:]):
                self.register_parameter('%dto%d_matrix' % (large, small), Parameter(torch.eye(small ** 2)))
        self.active_kernel_size = max(self.kernel_size_list)

    # -- which learned matrices feed the kernel for this ks ---------------------------------------
    def _matrices(self):
        """(m75, m53) as the C ABI wants them: `m75` is the 7->5 step, `m53` the final ->3 step
        (5->3, or a direct 7->3 when 5 is not in the list)."""
        def get(name):
            return getattr(self, name, None)
        kmax = max(self.kernel_size_list)
        m_big = get('7to5_matrix') if kmax == 7 else None
        m_small = get('5to3_matrix')
        if m_small is None:
            m_small = get('7to3_matrix')
        return m_big, m_small

    def _check_supported(self, kernel_size):
        if self.dilation != 1 or self.stride not in (1, 2, 3, 4):
            raise NotImplementedError('the B200 depthwise kernels cover dilation 1 and strides 1..4')
        if not set(self._ks_set) <= {3, 5, 7}:
            raise NotImplementedError('kernel sizes must come from {3, 5, 7}, got %s' % (self._ks_set,))
        if kernel_size not in self._ks_set and kernel_size != max(self.kernel_size_list):
            raise ValueError('kernel size %s is not in %s' % (kernel_size, self._ks_set))

    def get_active_filter(self, in_channel, kernel_size):
        """[in_channel, 1, ks, ks] filter the reference would hand to F.conv2d (dynamic_op.py:46-71)."""
        self._check_supported(kernel_size)
        m75, m53 = self._matrices()
        transform_on = self.KERNEL_TRANSFORM_MODE is not None
        w = self.conv.weight
        if not w.is_cuda:
            raise RuntimeError('libofa_sr_b200 has no CPU path: weights are on %s' % w.device)
        return OF.dw_active_filter(w.detach(), m75, m53, transform_on, kernel_size, in_channel)

    def forward(self, x, kernel_size=None):
        if kernel_size is None:
            kernel_size = self.active_kernel_size
        self._check_supported(kernel_size)
        get_same_padding(kernel_size)  # asserts an odd size, like the reference
        m75, m53 = self._matrices()
        transform_on = self.KERNEL_TRANSFORM_MODE is not None
        return OF.dw_conv(x, self.conv.weight, m75, m53, kernel_size, transform_on, self.stride)


class DynamicPointConv2d(nn.Module):

    def __init__(self, max_in_channels, max_out_channels, kernel_size=1, stride=1, dilation=1):
        super().__init__()
        self.max_in_channels = max_in_channels
        self.max_out_channels = max_out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.dilation = dilation
        self.conv = nn.Conv2d(
            self.max_in_channels, self.max_out_channels, self.kernel_size, stride=self.stride, bias=False,
        )
        self.active_out_channel = self.max_out_channels
        self._packed = OF.PackedWeightCache()

    def forward(self, x, out_channel=None):
        if out_channel is None:
            out_channel = self.active_out_channel
        if self.dilation != 1 and self.kernel_size != 1:
            raise NotImplementedError('the B200 conv kernels cover dilation 1 (every net of the reference uses 1)')
        in_channel = x.size(1)
        get_same_padding(self.kernel_size)
        s = self.stride
        if s != 1 and self.kernel_size == 1:
            x = x[:, :, ::s, ::s]                 # 1x1, padding 0: a strided conv reads every s-th pixel
            s = 1
        y = OF.conv2d(x, self.conv.weight, in_channel, out_channel, self.kernel_size)
        # k x k with "same" padding k // 2 (dynamic_op.py:110-111): the strided output is the stride-1 output at every
        # s-th position (the MobileNetV3-style first conv, ofa_mbv3.py:41-43; not on the SR nets' path)
        return y if s == 1 else y[:, :, ::s, ::s]


class DynamicLinear(nn.Module):
    """dynamic_op.py:115-136: F.linear(x, W[:out, :in], b[:out]) with the slice addressed in place."""

    def __init__(self, max_in_features, max_out_features, bias=True):
        super().__init__()
        self.max_in_features = max_in_features
        self.max_out_features = max_out_features
        self.bias = bias
        self.linear = nn.Linear(self.max_in_features, self.max_out_features, self.bias)
        self.active_out_features = self.max_out_features

    def forward(self, x, out_features=None):
        if out_features is None:
            out_features = self.active_out_features
        return OF.linear(x, self.linear.weight, self.linear.bias if self.bias else None, out_features)


class DynamicBatchNorm2d(nn.Module):
    SET_RUNNING_STATISTICS = False

    def __init__(self, max_feature_dim):
        super().__init__()
        self.max_feature_dim = max_feature_dim
        self.bn = nn.BatchNorm2d(self.max_feature_dim)

    @staticmethod
    def bn_forward(x, bn, feature_dim, act=B.ACT_NONE, residual=None):
        # A per-instance `forward` override is the hook elastic_nn.utils.set_running_statistics
        # installs on a deep copy (reference elastic_nn/utils.py:29-52); honour it.
        # (reference dynamic_op.py:150-151: `bn(x)` when C == num_features or SET_RUNNING_STATISTICS; OF.bn_act
        # calls the override when there is one and applies the fused activation / residual afterwards)
        return OF.bn_act(x, bn, feature_dim, act, residual)

    def forward(self, x):
        feature_dim = x.size(1)
        return self.bn_forward(x, self.bn, feature_dim)


class DynamicSE(SEModule):
    """dynamic_op.py:175-200: squeeze-and-excite on the active channel prefix; the reduce / expand 1x1 convs are sliced
    to [:num_mid, :C] and [:C, :num_mid] with num_mid = make_divisible(C // reduction, 8)."""

    def __init__(self, max_channel):
        super().__init__(max_channel)

    def forward(self, x):
        in_channel = x.size(1)
        num_mid = make_divisible(in_channel // self.reduction, divisor=8)
        return self._se(x, num_mid)
