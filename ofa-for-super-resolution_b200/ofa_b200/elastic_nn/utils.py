"""BatchNorm helpers shared by the elastic modules (reference ofa/elastic_nn/utils.py:69-82)."""
import torch


def adjust_bn_according_to_idx(bn, idx):
    """Permute a BatchNorm's per-channel tensors with `idx` (used by re_organize_middle_weights)."""
    for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var):
        t.data = torch.index_select(t.data, 0, idx)


def copy_bn(target_bn, src_bn):
    """Copy the first `target_bn.num_features` channels of `src_bn` (active-subnet extraction)."""
    c = target_bn.num_features
    target_bn.weight.data.copy_(src_bn.weight.data[:c])
    target_bn.bias.data.copy_(src_bn.bias.data[:c])
    target_bn.running_mean.data.copy_(src_bn.running_mean.data[:c])
    target_bn.running_var.data.copy_(src_bn.running_var.data[:c])
