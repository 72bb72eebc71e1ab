"""DynamicLinearLayer (reference: ofa/elastic_nn/modules/dynamic_layers.py:270-322) — the classifier head of the
MobileNetV3-style elastic nets: [dropout ->] DynamicLinear over the active input width.  Kept in its own module and
re-exported from dynamic_layers."""
import torch.nn as nn

from ...layers import LinearLayer
from ...utils import MyModule, get_net_device
from .dynamic_op import DynamicLinear

__all__ = ['DynamicLinearLayer']


class DynamicLinearLayer(MyModule):
    """Elastic classifier head: the input width follows whatever the previous layer produced."""

    def __init__(self, in_features_list, out_features, bias=True, dropout_rate=0):
        super().__init__()
        self.in_features_list, self.out_features = in_features_list, out_features
        self.bias, self.dropout_rate = bias, dropout_rate
        self.dropout = nn.Dropout(dropout_rate, inplace=True) if dropout_rate > 0 else None
        self.linear = DynamicLinear(max(in_features_list), out_features, bias)

    def forward(self, x):
        return self.linear(x if self.dropout is None else self.dropout(x))

    @property
    def module_str(self):
        return 'DyLinear(%d)' % self.out_features

    @property
    def config(self):
        # (sic) 'name' holds the OP's class name in the reference, and dropout_rate is not part of the config
        return dict(name=DynamicLinear.__name__, in_features_list=self.in_features_list,
                    out_features=self.out_features, bias=self.bias)

    @staticmethod
    def build_from_config(config):
        return DynamicLinearLayer(**config)

    def get_active_subnet(self, in_features, preserve_weight=True):
        """Static LinearLayer on the first `in_features` input columns."""
        sub = LinearLayer(in_features, self.out_features, self.bias, dropout_rate=self.dropout_rate)
        sub = sub.to(get_net_device(self))
        if preserve_weight:
            full = self.linear.linear
            sub.linear.weight.data.copy_(full.weight.data[:self.out_features, :in_features])
            if self.bias:
                sub.linear.bias.data.copy_(full.bias.data[:self.out_features])
        return sub
