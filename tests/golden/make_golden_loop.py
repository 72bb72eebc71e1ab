"""Golden for the drop-in proof: the UNMODIFIED reference training loop
(`ofa.elastic_nn.training.progressive_shrinking.train_one_epoch`, :94-224) driven for one epoch of three mini-batches on
the CPU, with the reference's RunConfig (cosine schedule, `build_optimizer` = torch.optim.Adam with the 'bn#bias'
no-decay split, sr_run_manager.py:67-133) and network (OFAMobileNetS4).  `dynamic_batch_size` = 2, seed rule :164.

    python tests/golden/make_golden_loop.py          ->  tests/golden/reference_loop.npz / .json

Only three things are patched, all outside the reference's files: `torch.Tensor.cuda` returns the tensor itself (the
loop calls `.cuda()` on every batch; this container has no GPU), `torch.Tensor.cpu` returns a detached copy (what a
device -> host copy is) and tqdm is silenced.  The run manager is a small
stand-in object with exactly the attributes the loop reads (net, optimizer, run_config, train_criterion); SRRunManager
itself needs data sets on disk.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import ofa_sr_oracle as O  # noqa: E402
from ofa.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d  # noqa: E402

DynamicSeparableConv2d.KERNEL_TRANSFORM_MODE = 1
from ofa.elastic_nn.networks.ofa_mbs4 import OFAMobileNetS4  # noqa: E402
from ofa.elastic_nn.training import progressive_shrinking as PS  # noqa: E402
from ofa.imagenet_codebase.run_manager.sr_run_manager import RunConfig  # noqa: E402

torch.set_num_threads(4)
FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])
WSEED, N_BATCH, BATCH, HR = 21, 3, 4, 32


def batches():
    rs = np.random.RandomState(77)
    out = []
    for _ in range(N_BATCH):
        hr = torch.from_numpy(rs.rand(BATCH, 3, HR, HR).astype(np.float32))
        out.append({'image': hr, '2x_down_image': torch.nn.functional.avg_pool2d(hr, 2),
                    '4x_down_image': torch.nn.functional.avg_pool2d(hr, 4)})
    return out


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self                      # no GPU here; the loop calls .cuda()
    # tensor2img_np clamps `tensor.float().cpu()` IN PLACE (sr_run_manager.py:567-597): harmless on a GPU, where .cpu()
    # copies -- on a CPU tensor it would edit the autograd graph's output; emulate the device -> host copy
    torch.Tensor.cpu = lambda self, *a, **k: self.detach().clone()
    PS.tqdm = lambda *a, **k: types.SimpleNamespace(__enter__=lambda s: s, __exit__=lambda s, *e: False,
                                                    set_postfix=lambda *a, **k: None, update=lambda *a, **k: None)

    class _T:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *e):
            return False

        def set_postfix(self, *a, **k):
            pass

        def update(self, *a, **k):
            pass
    PS.tqdm = _T

    net = OFAMobileNetS4(ks_list=list(FULL['ks_list']), expand_ratio_list=list(FULL['expand_ratio_list']),
                         depth_list=list(FULL['depth_list']), pixelshuffle_depth_list=[2])
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [2])
    net.load_state_dict(O.synth_state_dict(spec.param_shapes(), WSEED))

    class Cfg(RunConfig):
        def __init__(self):
            super().__init__(n_epochs=120, init_lr=1e-3, lr_schedule_type='cosine', lr_schedule_param=None,
                             dataset='none', train_batch_size=BATCH, test_batch_size=BATCH, valid_size=None,
                             opt_type='adam', opt_param=None, weight_decay=3e-5, label_smoothing=0,
                             no_decay_keys='bn#bias', mixup_alpha=None, model_init='he_fout', validation_frequency=1,
                             print_frequency=1)
            self._loader = batches()

        @property
        def train_loader(self):
            return self._loader
    cfg = Cfg()
    keys = cfg.no_decay_keys.split('#')
    params = [net.get_parameters(keys, mode='exclude'), net.get_parameters(keys, mode='include')]
    opt = cfg.build_optimizer([list(params[0]), list(params[1])])
    losses = []
    mse = torch.nn.MSELoss()

    def criterion(out, target):
        loss = mse(out, target)
        losses.append(float(loss.detach()))
        return loss
    rm = types.SimpleNamespace(net=net, optimizer=opt, run_config=cfg, train_criterion=criterion)
    args = types.SimpleNamespace(kd_ratio=0, dynamic_batch_size=2, independent_distributed_sampling=False, kd_type=None)
    mean_loss, mean_psnr = PS.train_one_epoch(rm, args, epoch=0, warmup_epochs=0, warmup_lr=0)

    sd = net.state_dict()
    out = {'mean_loss': np.float64(mean_loss), 'mean_psnr': np.float64(mean_psnr), 'losses': np.asarray(losses)}
    norms = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            norms[k] = float(v.double().norm())
    for k in ('dec_first_conv_block.conv.weight', 'blocks.0.mobile_inverted_conv.inverted_bottleneck.conv.conv.weight',
              'blocks.0.mobile_inverted_conv.depth_conv.conv.conv.weight', 'blocks.0.mobile_inverted_conv.depth_conv.conv.7to5_matrix',
              'blocks.3.mobile_inverted_conv.point_linear.conv.conv.weight', 'blocks.3.mobile_inverted_conv.depth_conv.bn.bn.running_mean',
              'blocks.3.mobile_inverted_conv.depth_conv.bn.bn.running_var', 'dec_final_output_conv_block.conv.weight',
              'dec_final_output_conv_block.bn.weight'):
        out['param/' + k] = sd[k].numpy()
    np.savez_compressed(os.path.join(HERE, 'reference_loop.npz'), **out)
    book = {'wseed': WSEED, 'n_batch': N_BATCH, 'batch': BATCH, 'hr': HR, 'init_lr': 1e-3, 'weight_decay': 3e-5,
            'n_epochs': 120, 'dynamic_batch_size': 2, 'norms': norms,
            'num_batches_tracked': {k: int(v) for k, v in sd.items() if k.endswith('num_batches_tracked')}}
    with open(os.path.join(HERE, 'reference_loop.json'), 'w') as f:
        json.dump(book, f)
    print('losses', losses, 'mean', mean_loss, 'psnr', mean_psnr)


if __name__ == '__main__':
    main()
