"""BatchNorm helpers shared by the elastic modules (reference ofa/elastic_nn/utils.py:16-82)."""
import copy

import torch
import torch.nn as nn

from .. import functional as OF
from ..utils import get_net_device


class _Avg:
    """AverageMeter / DistributedTensor of the reference (ofa/utils.py AverageMeter, imagenet_codebase/utils/
    __init__.py:119-140): batch-size-weighted running sum; `avg(world)` all-reduces the sum once when distributed
    (Horovod's allreduce averages over ranks)."""

    def __init__(self):
        self.sum, self.count = None, 0

    def update(self, val, n):
        self.sum = val * n if self.sum is None else self.sum + val * n
        self.count += n

    def avg(self, distributed, group=None):
        s = self.sum
        if distributed:
            import torch.distributed as dist
            s = s.clone()
            dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
            s = s / dist.get_world_size(group)
        return s / self.count


def set_running_statistics(model, data_loader, distributed=False, process_group=None):
    """BatchNorm re-calibration of the active sub-network (reference elastic_nn/utils.py:16-66): every BatchNorm2d
    of a deep copy gets a forward override that normalises with the statistics of the CURRENT batch and records
    them; afterwards the batch-size-weighted averages are written into the first C entries of the model's
    running_mean / running_var.  Statistics and normalisation run on the device (ofa_bn_stats / ofa_affine_act);
    with `distributed` the per-rank sums are all-reduced once per BatchNorm (NCCL)."""
    from .modules.dynamic_op import DynamicBatchNorm2d
    bn_mean, bn_var = {}, {}
    forward_model = copy.deepcopy(model)
    for name, m in forward_model.named_modules():
        if isinstance(m, nn.BatchNorm2d):
            bn_mean[name], bn_var[name] = _Avg(), _Avg()

            def new_forward(bn, mean_est, var_est):
                def lambda_forward(x):
                    batch_mean, batch_var = OF.batch_stats(x)          # biased variance, as the reference computes
                    mean_est.update(batch_mean, x.size(0))
                    var_est.update(batch_var, x.size(0))
                    return OF.normalize_with(x, batch_mean, batch_var, bn.weight, bn.bias, bn.eps)
                return lambda_forward
            m.forward = new_forward(m, bn_mean[name], bn_var[name])
    with torch.no_grad():
        DynamicBatchNorm2d.SET_RUNNING_STATISTICS = True
        try:
            for images in data_loader:
                images = images['image'].to(get_net_device(forward_model))
                forward_model(images)
        finally:
            DynamicBatchNorm2d.SET_RUNNING_STATISTICS = False
    for name, m in model.named_modules():
        if name in bn_mean and bn_mean[name].count > 0:
            mean = bn_mean[name].avg(distributed, process_group)
            var = bn_var[name].avg(distributed, process_group)
            feature_dim = mean.size(0)
            assert isinstance(m, nn.BatchNorm2d)
            m.running_mean.data[:feature_dim].copy_(mean)
            m.running_var.data[:feature_dim].copy_(var)


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
