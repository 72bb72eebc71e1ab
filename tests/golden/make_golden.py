"""Generates tests/golden/*.npz / *.json by running the UNMODIFIED reference (imported from
/root/reference, CPU, fp32).  Run once in the build container:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these fixtures are the pin of the
oracle (oracle/ofa_sr_oracle.py) and, through it, of the CUDA path.  Weights come from the oracle's
`synth_state_dict` recipe (numpy RandomState — identical on every machine), loaded into the
reference networks with `load_state_dict`; only inputs, outputs and summaries are stored.
"""
import json
import os
import random
import sys

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
from ofa.elastic_nn.networks.ofa_mbx4 import OFAMobileNetX4  # noqa: E402
from ofa.elastic_nn.modules.dynamic_layers import DynamicMBConvLayer  # noqa: E402

torch.set_num_threads(4)
FULL = dict(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4])

NET_CASES = {
    # name: (kind, pixelshuffle_depth_list, input shape, weight seed)
    's4_ps1': ('s4', [1], (1, 3, 16, 12), 11),
    's4_ps12': ('s4', [1, 2], (2, 3, 12, 16), 12),
    'x4_ps12': ('x4', [1, 2], (1, 3, 16, 24), 13),
}
SUBNETS = [
    dict(ks=3, e=3, d=2, pixel_d=1),
    dict(ks=7, e=6, d=4, pixel_d=2),
    dict(ks=5, e=4, d=3, pixel_d=1),
    'sample:0', 'sample:1', 'sample:2', 'sample:3',
]


def build_ref(kind, pd):
    cls = OFAMobileNetS4 if kind == 's4' else OFAMobileNetX4
    return cls(ks_list=list(FULL['ks_list']), expand_ratio_list=list(FULL['expand_ratio_list']),
               depth_list=list(FULL['depth_list']), pixelshuffle_depth_list=list(pd))


def rand_image(shape, seed):
    return torch.from_numpy(np.random.RandomState(seed).rand(*shape).astype(np.float32))


def main():
    out = {}
    book = {}
    for name, (kind, pd, shape, wseed) in NET_CASES.items():
        net = build_ref(kind, pd)
        spec = O.SuperNetSpec(kind, FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], pd)
        shapes = spec.param_shapes()
        ref_sd = net.state_dict()
        assert list(ref_sd.keys()) == list(shapes.keys()), 'oracle parameter inventory != reference state_dict'
        assert all(tuple(ref_sd[k].shape) == tuple(shapes[k]) for k in shapes)
        net.load_state_dict(O.synth_state_dict(shapes, wseed))
        net.eval()
        x = rand_image(shape, 100 + wseed)
        out[name + '/x'] = x.numpy()
        book[name] = {'kind': kind, 'pd': pd, 'wseed': wseed, 'keys': list(shapes.keys()), 'subnets': []}
        for i, sub in enumerate(SUBNETS):
            if isinstance(sub, str):
                random.seed(int(sub.split(':')[1]))
                setting = net.sample_active_subnet()
            else:
                net.set_active_subnet(**sub)
                setting = dict(sub)
            with torch.no_grad():
                y = net(x)
            out['%s/y%d' % (name, i)] = y.numpy()
            book[name]['subnets'].append({
                'request': sub, 'setting': setting, 'runtime_depth': list(net.runtime_depth),
                'active': [[b.mobile_inverted_conv.active_kernel_size, b.mobile_inverted_conv.active_expand_ratio]
                           for b in net.blocks if hasattr(b, 'mobile_inverted_conv')],
                'out_shape': list(y.shape),
            })

    # ---- sampling stream: 40 seeds per net (bit-exact subnet selection, Q6) -------------------------
    for kind, pd in (('s4', [1, 2]), ('x4', [1, 2]), ('s4', [1])):
        net = build_ref(kind, pd)
        rows = []
        for seed in range(40):
            random.seed(seed)
            s = net.sample_active_subnet()
            rows.append({'seed': seed, 'setting': s, 'runtime_depth': list(net.runtime_depth)})
        # constraint path (progressive shrinking narrows the candidate lists)
        net.set_constraint([5, 7], 'kernel_size')
        net.set_constraint([3, 4], 'depth')
        random.seed(123)
        s = net.sample_active_subnet()
        rows.append({'seed': 123, 'constraint': {'kernel_size': [5, 7], 'depth': [3, 4]}, 'setting': s,
                     'runtime_depth': list(net.runtime_depth)})
        book['sampling_%s_ps%s' % (kind, ''.join(map(str, pd)))] = rows

    # ---- single elastic MBConv block: eval outputs, active filters, one training step -------------
    torch.manual_seed(0)
    layer = DynamicMBConvLayer([64], [64], kernel_size_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], stride=1,
                               act_func='relu6', use_se=False)
    spec = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1])
    blk = {k[len('blocks.0.mobile_inverted_conv.'):]: v for k, v in spec.param_shapes().items()
           if k.startswith('blocks.0.mobile_inverted_conv.')}
    layer.load_state_dict(O.synth_state_dict(blk, 21))
    x = torch.from_numpy(np.random.RandomState(5).randn(2, 64, 9, 11).astype(np.float32))
    out['block/x'] = x.numpy()
    for ks in (3, 5, 7):
        for e in (3, 4, 6):
            layer.active_kernel_size, layer.active_expand_ratio = ks, e
            layer.eval()
            with torch.no_grad():
                out['block/eval_k%d_e%d' % (ks, e)] = layer(x).numpy()
        mid = 384
        out['block/filter_k%d' % ks] = layer.depth_conv.conv.get_active_filter(mid, ks).detach().numpy()
    # training step (a5 training BN, a14 backward) for (ks=3, e=4) and (ks=5, e=6)
    for ks, e in ((3, 4), (5, 6), (7, 3)):
        layer.load_state_dict(O.synth_state_dict(blk, 21))
        layer.train()
        layer.zero_grad()
        layer.active_kernel_size, layer.active_expand_ratio = ks, e
        xin = x.clone().requires_grad_(True)
        y = layer(xin)
        tgt = torch.from_numpy(np.random.RandomState(6).randn(*y.shape).astype(np.float32))
        loss = torch.nn.functional.mse_loss(y, tgt)
        loss.backward()
        tag = 'block/train_k%d_e%d/' % (ks, e)
        out[tag + 'y'] = y.detach().numpy()
        out[tag + 'target'] = tgt.numpy()
        out[tag + 'loss'] = np.float32(loss.item())
        out[tag + 'dx'] = xin.grad.numpy()
        for pname, p in layer.named_parameters():
            out[tag + 'grad/' + pname] = (p.grad.numpy() if p.grad is not None else np.zeros(0, np.float32))
        for bname, b in layer.named_buffers():
            out[tag + 'buf/' + bname] = b.numpy().copy()

    # ---- one S4 progressive-shrinking style training step (two sampled subnets, grads accumulate) ----
    net = build_ref('s4', [1, 2])
    shapes = O.SuperNetSpec('s4', FULL['ks_list'], FULL['expand_ratio_list'], FULL['depth_list'], [1, 2]).param_shapes()
    net.load_state_dict(O.synth_state_dict(shapes, 31))
    net.train()
    lr_img = rand_image((2, 3, 8, 8), 77)
    hr_img = rand_image((2, 3, 32, 32), 78)
    out['train_s4/lr'] = lr_img.numpy()
    out['train_s4/hr'] = hr_img.numpy()
    net.zero_grad()
    losses = []
    for j in range(2):
        random.seed(int('%d%.3d%.3d' % (5, j, 0)))   # progressive_shrinking.py:164
        net.sample_active_subnet()
        y = net(lr_img)
        loss = torch.nn.functional.mse_loss(y, hr_img)
        loss.backward()
        losses.append(loss.item())
    out['train_s4/losses'] = np.asarray(losses, np.float32)
    gnorm = {}
    for pname, p in net.named_parameters():
        gnorm[pname] = float(p.grad.norm().item()) if p.grad is not None else None
    book['train_s4_grad_norms'] = gnorm
    out['train_s4/grad/dec_first_conv_block.conv.weight'] = net.dec_first_conv_block.conv.weight.grad.numpy()
    out['train_s4/grad/blocks.0.mobile_inverted_conv.depth_conv.conv.conv.weight'] = \
        net.blocks[0].mobile_inverted_conv.depth_conv.conv.conv.weight.grad.numpy()
    out['train_s4/buf/blocks.0.mobile_inverted_conv.depth_conv.bn.bn.running_var'] = \
        net.blocks[0].mobile_inverted_conv.depth_conv.bn.bn.running_var.numpy().copy()
    book['train_s4_runtime_depth'] = list(net.runtime_depth)

    np.savez_compressed(os.path.join(HERE, 'reference_outputs.npz'), **out)
    with open(os.path.join(HERE, 'reference_bookkeeping.json'), 'w') as f:
        json.dump(book, f, indent=1, sort_keys=True)
    print('wrote', len(out), 'arrays;', os.path.getsize(os.path.join(HERE, 'reference_outputs.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
