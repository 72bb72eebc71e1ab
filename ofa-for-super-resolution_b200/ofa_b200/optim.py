"""Optimizer and learning-rate schedule of the SR run manager.

`FusedAdam` is torch.optim.Adam as sr_run_manager.py:115-133,180-185 builds it — two parameter groups, L2 weight
decay on everything except the `no_decay_keys` ('bn#bias') — executed as ONE multi-tensor kernel launch per step
(plus a 1-block counter bump) instead of a few launches per parameter tensor.  Parameters whose .grad is None (blocks
outside the sampled sub-network) are skipped exactly as optimizer.step() skips them.
`cosine_lr` / `warmup_lr` restate RunConfig.calc_learning_rate / warmup_adjust_learning_rate (:67-90).
"""
import math

import numpy as np
import torch

from . import backend as B

_CHUNK = 2048


def cosine_lr(init_lr, n_epochs, epoch, batch=0, n_batch=None):
    """sr_run_manager.py:67-76, lr_schedule_type == 'cosine'."""
    t_total = n_epochs * n_batch
    t_cur = epoch * n_batch + batch
    return 0.5 * init_lr * (1 + math.cos(math.pi * t_cur / t_total))


def warmup_lr(init_lr, t_total, n_batch, epoch, batch=0, warmup_lr=0):
    """sr_run_manager.py:85-90."""
    t_cur = epoch * n_batch + batch + 1
    return t_cur / t_total * (init_lr - warmup_lr) + warmup_lr


def split_no_decay(named_parameters, no_decay_keys='bn#bias'):
    """network.get_parameters(keys, mode='exclude' / 'include') (networks' get_parameters): (decay, no_decay)."""
    keys = no_decay_keys.split('#') if no_decay_keys else []
    decay, no_decay = [], []
    for name, p in named_parameters:
        (no_decay if any(k in name for k in keys) else decay).append(p)
    return decay, no_decay


class FusedAdam:
    def __init__(self, decay_params, no_decay_params=(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        groups = [(p, float(weight_decay)) for p in decay_params if p.requires_grad] + \
                 [(p, 0.0) for p in no_decay_params if p.requires_grad]
        assert groups, 'no parameters'
        self.params = [p for p, _ in groups]
        self.betas, self.eps = betas, eps
        # torch.optim-style view: the run manager writes param_group['lr'] (sr_run_manager.py:78-90) and checkpoints
        # optimizer.state_dict() (sr_run_manager.py:301,539; progressive_shrinking.py:252)
        self.param_groups = [
            {'params': [p for p in decay_params if p.requires_grad], 'lr': lr, 'betas': betas, 'eps': eps,
             'weight_decay': float(weight_decay)},
            {'params': [p for p in no_decay_params if p.requires_grad], 'lr': lr, 'betas': betas, 'eps': eps,
             'weight_decay': 0.0},
        ]
        dev = self.params[0].device
        assert all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self.params), \
            'FusedAdam updates contiguous fp32 CUDA parameters in place'
        total = sum(p.numel() for p in self.params)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        # descriptor table (OfaAdamTensor: 3 pointers, int64 numel, float wd, int32 pad = 40 bytes) and chunk map
        table = np.zeros(len(groups), dtype=[('p', '<u8'), ('m', '<u8'), ('v', '<u8'), ('numel', '<i8'), ('wd', '<f4'), ('pad', '<i4')])
        chunks, off = [], 0
        for i, (p, wd) in enumerate(groups):
            table[i] = (p.data_ptr(), self.exp_avg.data_ptr() + 4 * off, self.exp_avg_sq.data_ptr() + 4 * off, p.numel(), wd, 0)
            chunks += [(i, c) for c in range((p.numel() + _CHUNK - 1) // _CHUNK)]
            off += p.numel()
        self._table = torch.from_numpy(table.view(np.uint8).copy()).to(dev)
        self._chunks = torch.tensor(chunks, dtype=torch.int32, device=dev)
        self._steps = torch.zeros(len(groups), dtype=torch.int32, device=dev)
        self._grad_dev = torch.zeros(len(groups), dtype=torch.int64, device=dev)
        self._ptrs = [p.data_ptr() for p in self.params]

    @property
    def lr(self):
        return self.param_groups[0]['lr']

    @lr.setter
    def lr(self, value):
        for g in self.param_groups:
            g['lr'] = value

    def state_dict(self):
        """Moments, per-tensor step counters and hyper-parameters (a resumed run continues Adam's bias correction)."""
        return {'state': {'exp_avg': self.exp_avg.clone(), 'exp_avg_sq': self.exp_avg_sq.clone(),
                          'steps': self._steps.clone()},
                'param_groups': [{k: v for k, v in g.items() if k != 'params'} for g in self.param_groups]}

    def load_state_dict(self, sd):
        st = sd['state']
        assert st['exp_avg'].numel() == self.exp_avg.numel(), 'optimizer state belongs to another parameter set'
        self.exp_avg.copy_(st['exp_avg'])
        self.exp_avg_sq.copy_(st['exp_avg_sq'])
        self._steps.copy_(st['steps'])
        for g, saved in zip(self.param_groups, sd['param_groups']):
            g.update({k: v for k, v in saved.items() if k != 'weight_decay'})
        self.betas, self.eps = tuple(self.param_groups[0]['betas']), self.param_groups[0]['eps']

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    _RING = 4

    def step(self):
        assert [p.data_ptr() for p in self.params] == self._ptrs, 'parameters were re-allocated: rebuild FusedAdam'
        ptrs, touched = [], []
        for p in self.params:
            if p.grad is None:
                ptrs.append(0)
            else:
                assert p.grad.dtype == torch.float32 and p.grad.is_contiguous()
                ptrs.append(p.grad.data_ptr())
                touched.append(p)
        if torch.cuda.is_current_stream_capturing():
            # CUDA-graph capture of a whole training step: the gradient buffers live in the graph's private pool, so
            # their addresses are the same on every replay; the table is uploaded from a pinned buffer that is not
            # touched again after capture
            self._grad_pinned = torch.tensor(ptrs, dtype=torch.int64).pin_memory()
            self._grad_dev.copy_(self._grad_pinned, non_blocking=True)
        else:
            # a blocking copy from pageable memory synchronises the stream: the host could never run ahead of the GPU
            # into the next step.  Upload from a small ring of pinned buffers instead; a slot is reused only after the
            # copy that last read it has completed (it is _RING steps old by then).
            ring = self.__dict__.setdefault('_grad_ring', [])
            if len(ring) < self._RING:
                ring.append([torch.empty(len(ptrs), dtype=torch.int64).pin_memory(), None])
                slot = ring[-1]
                self._ring_pos = len(ring) % self._RING
            else:
                slot = ring[self._ring_pos]
                self._ring_pos = (self._ring_pos + 1) % self._RING
                if slot[1] is not None:
                    slot[1].synchronize()
            slot[0].copy_(torch.tensor(ptrs, dtype=torch.int64))
            self._grad_dev.copy_(slot[0], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self._grad_dev.device))
            slot[1] = ev
        dev = self.params[0].device
        B.check(B.lib().ofa_adam_step(self._table.data_ptr(), self._chunks.data_ptr(), len(self.params),
                                      self._chunks.shape[0], self._grad_dev.data_ptr(), self._steps.data_ptr(),
                                      float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                      B.stream_ptr(dev)))
        # the kernel writes the parameters through raw pointers: tell autograd / the modules' derived 16-bit weight
        # caches (keyed on Tensor._version) that these tensors changed, as an in-place torch op would have
        if touched:
            torch.autograd.graph.increment_version(touched)

    def set_lr(self, lr):
        """adjust_learning_rate / warmup_adjust_learning_rate write param_group['lr'] (sr_run_manager.py:78-90)."""
        self.lr = lr
