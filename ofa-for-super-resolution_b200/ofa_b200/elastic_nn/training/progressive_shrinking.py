"""The progressive-shrinking training iteration of the SR supernets on the B200 path.

Restates the inner loop of the reference's `train_one_epoch` (ofa/elastic_nn/training/progressive_shrinking.py:126-203)
around the drop-in modules: per mini-batch, `dynamic_batch_size` sub-networks are sampled -- each after seeding Python's
`random` with the reference's rule (:164) so that every data-parallel rank draws the SAME sub-network --, each runs
forward + loss + backward on the 2x- or 4x-downscaled input its `pixel_d` selects (:176-180), gradients accumulate over
the sub-networks, then ONE optimizer step.  What is not rebuilt: the run manager, tqdm / logging, the data loader
(out of scope).  Differences, all on the host side:
  * the per-step PSNR (reference :196 -- a device -> host copy of the whole batch every sub-network) is computed ON the
    device by `ofa_b200.metrics.psnr_y` (bit-exact, tests/golden/reference_metric.json) and only when asked for;
  * data-parallel training is one process per GPU: after the last backward the gradients are averaged with ONE flat
    all-reduce (`parallel.FlatGradAllReduce`, Horovod semantics of distributed_run_manager.py:72-75).
"""
import random

import torch
import torch.nn.functional as F

__all__ = ['subnet_seed', 'train_step', 'train_one_epoch']


def subnet_seed(epoch, n_batch, i, k):
    """progressive_shrinking.py:164 -- int('%d%.3d%.3d' % (epoch * nBatch + i, k, 0))."""
    return int('%d%.3d%.3d' % (epoch * n_batch + i, k, 0))


def train_step(net, optimizer, mini_batch, epoch=0, i=0, n_batch=1, dynamic_batch_size=1, criterion=None, reducer=None,
               lr=None, teacher=None, kd_ratio=0.0, psnr=False, boundary_module=None):
    """One mini-batch.  `mini_batch` = {'image', '2x_down_image', '4x_down_image'} device tensors (div2k_setxx.py:166-171
    keys).  Returns (mean loss over the sub-networks [0-dim device tensor], list of sampled settings, mean PSNR or None).
    `teacher` + `kd_ratio` > 0: the `kd_type != 'ce'` branch (:189-192): loss = (kd * MSE(out, teacher_out) +
    criterion(out, target)) * 2 / (kd + 1), teacher forward under no_grad in train mode (:146-149)."""
    criterion = criterion or F.mse_loss
    images = mini_batch['image']
    if lr is not None:
        if hasattr(optimizer, 'set_lr'):
            optimizer.set_lr(lr)
        else:
            for g in optimizer.param_groups:
                g['lr'] = lr
    soft = None
    if kd_ratio > 0:
        teacher.train()
        with torch.no_grad():
            soft = teacher(images).detach()
    optimizer.zero_grad(set_to_none=True) if _accepts_set_to_none(optimizer) else optimizer.zero_grad()
    losses, settings_all, psnrs = [], [], []
    for k in range(dynamic_batch_size):
        random.seed(subnet_seed(epoch, n_batch, i, k))
        settings = net.sample_active_subnet()
        settings_all.append(settings)
        pixel_d = settings['pixel_d'][0]
        x = mini_batch['2x_down_image'] if pixel_d == 1 else mini_batch['4x_down_image']
        if reducer is not None and boundary_module is not None and k == dynamic_batch_size - 1:
            _arm_overlap(reducer, boundary_module)
        out = net(x)
        if kd_ratio == 0:
            loss = criterion(out, images)
        else:
            loss = (kd_ratio * F.mse_loss(out, soft) + criterion(out, images)) * (2 / (kd_ratio + 1))
        if psnr:
            from ... import metrics
            psnrs.append(metrics.psnr_y(out.detach(), images))
        losses.append(loss.detach())
        loss.backward()
    if reducer is not None:
        reducer.reduce()
    optimizer.step()
    mean_loss = torch.stack(losses).mean()
    return mean_loss, settings_all, (sum(psnrs) / len(psnrs) if psnrs else None)


def _accepts_set_to_none(optimizer):
    return True


def _arm_overlap(reducer, boundary_module):
    """The tail segment's all-reduce starts as soon as the LAST backward of the step reaches the input of the tail layers."""
    handle = {}

    def pre_hook(module, inputs):
        if inputs and torch.is_tensor(inputs[0]) and inputs[0].requires_grad:
            reducer.watch(inputs[0])
        handle['h'].remove()
    handle['h'] = boundary_module.register_forward_pre_hook(pre_hook)


def train_one_epoch(net, optimizer, loader, epoch, lr_schedule=None, dynamic_batch_size=1, reducer=None, teacher=None,
                    kd_ratio=0.0, psnr=False, boundary_module=None, to_device=None):
    """train_one_epoch (:94-224) without the run manager: `loader` yields the reference's mini-batch dicts; `lr_schedule(
    epoch, i, n_batch)` -> learning rate (sr_run_manager.py:67-90 through ofa_b200.optim.cosine_lr / warmup_lr).  Returns
    (mean loss, mean PSNR or None) as Python floats -- ONE device -> host read at the end of the epoch."""
    net.train()
    n_batch = len(loader)
    loss_sum = None
    psnr_sum, count = 0.0, 0
    for i, mini_batch in enumerate(loader):
        if to_device is not None:
            mini_batch = {k: (v.to(to_device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in mini_batch.items()}
        lr = lr_schedule(epoch, i, n_batch) if lr_schedule is not None else None
        loss, _, p = train_step(net, optimizer, mini_batch, epoch, i, n_batch, dynamic_batch_size, reducer=reducer, lr=lr,
                                teacher=teacher, kd_ratio=kd_ratio, psnr=psnr, boundary_module=boundary_module)
        n = mini_batch['image'].size(0)
        loss_sum = loss * n if loss_sum is None else loss_sum + loss * n
        if p is not None:
            psnr_sum += float(p) * n
        count += n
    mean_loss = float(loss_sum / count) if count else 0.0
    return mean_loss, (psnr_sum / count if psnr and count else None)
