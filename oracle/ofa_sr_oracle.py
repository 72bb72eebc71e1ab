"""CPU ORACLE of the elastic-MBConv super-resolution hot path — TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, what the reference (twice154/ofa-for-super-resolution) computes on
the path of SURVEY.md §8.  It is imported only by tests/, by __graft_entry__.smoke() and by the
`cpu_baseline` / `--impl reference` legs of bench.py, always as the CHECKER or the BASELINE — never by
the product package `ofa_b200`, which has no CPU path at all.

The reference's arithmetic is PyTorch's own CPU kernels (F.conv2d / F.batch_norm / F.linear /
pixel_shuffle: SURVEY §8c "third-party arithmetic"), so the restatement uses the same torch CPU
functional ops on plain tensors — no nn.Module, no product code.  Every function cites the
reference lines it follows (paths relative to the reference tree).

Pinned: tests/golden/*.npz were produced by importing the UNMODIFIED reference in the build
container (tests/golden/make_golden.py); tests/test_oracle_golden.py checks this file against them.
The reference itself ships no tests or golden vectors.
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------
# integer helpers (bit-exact)
# ---------------------------------------------------------------------------------------------

def make_divisible(v, divisor, min_val=None):
    """ofa/imagenet_codebase/utils/pytorch_modules.py:12-29"""
    if min_val is None:
        min_val = divisor
    new_v = max(min_val, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def sub_filter_start_end(kernel_size, sub_kernel_size):
    """ofa/imagenet_codebase/utils/__init__.py:84-89"""
    center = kernel_size // 2
    dev = sub_kernel_size // 2
    return center - dev, center + dev + 1


# ---------------------------------------------------------------------------------------------
# ops
# ---------------------------------------------------------------------------------------------

def active_filter(w, mats, ks_set, C, ks, transform_on=True):
    """DynamicSeparableConv2d.get_active_filter — dynamic_op.py:46-71.
    w: [Cmax,1,kmax,kmax]; mats: {'7to5_matrix': [25,25], '5to3_matrix': [9,9]}; ks_set sorted."""
    kmax = w.shape[-1]
    s, e = sub_filter_start_end(kmax, ks)
    filters = w[:C, :, s:e, s:e]
    if transform_on and ks < kmax:
        cur = w[:C]
        for i in range(len(ks_set) - 1, 0, -1):
            src = ks_set[i]
            if src <= ks:
                break
            tgt = ks_set[i - 1]
            s, e = sub_filter_start_end(src, tgt)
            flat = cur[:, :, s:e, s:e].contiguous().view(-1, tgt * tgt)
            flat = F.linear(flat, mats['%dto%d_matrix' % (src, tgt)])
            cur = flat.view(C, 1, tgt, tgt)
        filters = cur
    return filters.contiguous()


def dw_conv(x, filt):
    """DynamicSeparableConv2d.forward — dynamic_op.py:73-84 (stride 1, same padding, groups = C)."""
    ks = filt.shape[-1]
    return F.conv2d(x, filt, None, 1, ks // 2, 1, x.shape[1])


def sliced_conv(x, w, cout):
    """DynamicPointConv2d.forward — dynamic_op.py:104-112 (also the static k x k convs of ConvLayer,
    layers.py:135-147, with cout = all)."""
    cin = x.shape[1]
    ks = w.shape[-1]
    return F.conv2d(x, w[:cout, :cin].contiguous(), None, 1, ks // 2, 1, 1)


_RECAL_SINK = None


def batch_norm(x, sd, prefix, training=False, momentum=0.1, eps=1e-5):
    """DynamicBatchNorm2d.bn_forward — dynamic_op.py:148-167.  Running statistics in `sd` are updated
    in place on the [:C] slice when training, and num_batches_tracked is bumped."""
    C = x.shape[1]
    rm, rv = sd[prefix + 'running_mean'], sd[prefix + 'running_var']
    if _RECAL_SINK is not None:
        # set_running_statistics (elastic_nn/utils.py:29-47): normalise with THIS batch's statistics (biased
        # variance) and record them, weighted by the batch size
        mean = x.mean(0, keepdim=True).mean(2, keepdim=True).mean(3, keepdim=True)
        var = ((x - mean) * (x - mean)).mean(0, keepdim=True).mean(2, keepdim=True).mean(3, keepdim=True)
        mean, var = torch.squeeze(mean), torch.squeeze(var)
        _RECAL_SINK.append((prefix, mean.reshape(-1), var.reshape(-1), x.shape[0]))
        return F.batch_norm(x, mean.reshape(-1), var.reshape(-1), sd[prefix + 'weight'][:C], sd[prefix + 'bias'][:C],
                            False, 0.0, eps)
    if training:
        sd[prefix + 'num_batches_tracked'] += 1
    return F.batch_norm(x, rm[:C], rv[:C], sd[prefix + 'weight'][:C], sd[prefix + 'bias'][:C], training,
                        momentum if training else 0.0, eps)


def relu6(x):
    """build_activation('relu6') — ofa/utils.py:245-246"""
    return torch.clamp(x, 0.0, 6.0)


def pixel_shuffle2(x):
    """nn.PixelShuffle(2) — ofa/utils.py:259-260: out[n,c,2h+i,2w+j] = in[n,4c+2i+j,h,w]"""
    n, c, h, w = x.shape
    return x.view(n, c // 4, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3).reshape(n, c // 4, 2 * h, 2 * w)


def pixel_unshuffle2(x):
    """pixel_unshuffle — ofa/utils.py:383-397: out[n,4c+2y+x,h,w] = in[n,c,2h+y,2w+x]"""
    n, c, h, w = x.shape
    return x.view(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(n, 4 * c, h // 2, w // 2)


def conv_layer(x, sd, prefix, act=None, training=False, momentum=0.1, eps=1e-5):
    """ConvLayer = conv -> BN -> act (My2DLayer.forward, layers.py:94-98)."""
    w = sd[prefix + 'conv.weight']
    y = sliced_conv(x, w, w.shape[0])
    y = batch_norm(y, sd, prefix + 'bn.', training, momentum, eps)
    if act == 'pixelshuffle':
        y = pixel_shuffle2(y)
    elif act == 'pixelunshuffle':
        y = pixel_unshuffle2(y)
    elif act == 'relu6':
        y = relu6(y)
    return y


def mbconv_block(x, sd, prefix, ks, expand, ks_set, training=False, momentum=0.1, eps=1e-5, transform_on=True):
    """MobileInvertedResidualBlock(DynamicMBConvLayer, IdentityLayer) — dynamic_layers.py:70-84 +
    proxyless_nets.py:44-51.  prefix e.g. 'blocks.3.'"""
    p = prefix + 'mobile_inverted_conv.'
    cin = x.shape[1]
    mid = make_divisible(round(cin * expand), 8)
    h = sliced_conv(x, sd[p + 'inverted_bottleneck.conv.conv.weight'], mid)
    h = relu6(batch_norm(h, sd, p + 'inverted_bottleneck.bn.bn.', training, momentum, eps))
    mats = {k.split('.')[-1]: v for k, v in sd.items() if k.startswith(p + 'depth_conv.conv.') and k.endswith('_matrix')}
    filt = active_filter(sd[p + 'depth_conv.conv.conv.weight'], mats, ks_set, mid, ks, transform_on)
    h = relu6(batch_norm(dw_conv(h, filt), sd, p + 'depth_conv.bn.bn.', training, momentum, eps))
    cout = sd[p + 'point_linear.conv.conv.weight'].shape[0]
    h = batch_norm(sliced_conv(h, sd[p + 'point_linear.conv.conv.weight'], cout), sd, p + 'point_linear.bn.bn.',
                   training, momentum, eps)
    return h + x


# ---------------------------------------------------------------------------------------------
# the MobileNetV3 flavour of the elastic modules: SE, h-swish, stride, linear
# ---------------------------------------------------------------------------------------------

def h_sigmoid(x):
    """ofa/utils.py:345-352 Hsigmoid"""
    return F.relu6(x + 3.0) / 6.0


def h_swish(x):
    """ofa/utils.py:334-341 Hswish"""
    return x * F.relu6(x + 3.0) / 6.0


def apply_act(x, act_func):
    """build_activation vocabulary used by the elastic blocks — ofa/utils.py:242-257"""
    if act_func is None:
        return x
    return {'relu6': relu6, 'relu': F.relu, 'h_swish': h_swish, 'h_sigmoid': h_sigmoid}[act_func](x)


def dynamic_se(x, sd, prefix, reduction=4):
    """DynamicSE.forward — dynamic_op.py:175-200 (SEModule: ofa/utils.py:354-375).  prefix e.g. '...depth_conv.se.'"""
    C = x.shape[1]
    num_mid = make_divisible(C // reduction, divisor=8)
    y = x.mean(3, keepdim=True).mean(2, keepdim=True)
    y = F.relu(F.conv2d(y, sd[prefix + 'fc.reduce.weight'][:num_mid, :C], sd[prefix + 'fc.reduce.bias'][:num_mid]))
    y = h_sigmoid(F.conv2d(y, sd[prefix + 'fc.expand.weight'][:C, :num_mid], sd[prefix + 'fc.expand.bias'][:C]))
    return x * y


def dynamic_linear(x, weight, bias, out_features):
    """DynamicLinear.forward — dynamic_op.py:128-136"""
    return F.linear(x, weight[:out_features, :x.shape[1]], bias[:out_features] if bias is not None else None)


def dynamic_mbconv(x, sd, prefix, ks, expand, out_channel, ks_set, stride=1, act_func='relu6', use_se=False,
                   training=False, momentum=0.1, eps=1e-5, transform_on=True):
    """DynamicMBConvLayer.forward in its general form — dynamic_layers.py:14-84: sliced expand -> BN -> act ->
    elastic depthwise with stride (same padding ks // 2, dynamic_op.py:73-84) -> BN -> act [-> DynamicSE] -> sliced
    project to `out_channel` -> BN.  No residual (MobileInvertedResidualBlock adds it only when shapes agree)."""
    cin = x.shape[1]
    mid = make_divisible(round(cin * expand), 8)
    h = x
    if (prefix + 'inverted_bottleneck.conv.conv.weight') in sd:
        h = sliced_conv(x, sd[prefix + 'inverted_bottleneck.conv.conv.weight'], mid)
        h = apply_act(batch_norm(h, sd, prefix + 'inverted_bottleneck.bn.bn.', training, momentum, eps), act_func)
    else:
        mid = cin
    mats = {k.split('.')[-1]: v for k, v in sd.items()
            if k.startswith(prefix + 'depth_conv.conv.') and k.endswith('_matrix')}
    filt = active_filter(sd[prefix + 'depth_conv.conv.conv.weight'], mats, ks_set, mid, ks, transform_on)
    h = F.conv2d(h, filt, None, stride, ks // 2, 1, mid)
    h = apply_act(batch_norm(h, sd, prefix + 'depth_conv.bn.bn.', training, momentum, eps), act_func)
    if use_se:
        h = dynamic_se(h, sd, prefix + 'depth_conv.se.')
    h = sliced_conv(h, sd[prefix + 'point_linear.conv.conv.weight'], out_channel)
    return batch_norm(h, sd, prefix + 'point_linear.bn.bn.', training, momentum, eps)


# ---------------------------------------------------------------------------------------------
# supernets: topology + sub-network bookkeeping (ofa_mbs4.py:20-178,263-370; ofa_mbx4.py:20-254,345-453)
# ---------------------------------------------------------------------------------------------

class SuperNetSpec:
    """Architecture + active-subnet state of OFAMobileNetS4 ('s4') or OFAMobileNetX4 ('x4')."""

    def __init__(self, kind, ks_list, expand_ratio_list, depth_list, pixelshuffle_depth_list):
        assert kind in ('s4', 'x4')
        self.kind = kind
        self.ks_list = sorted(ks_list)
        self.expand_ratio_list = sorted(expand_ratio_list)
        self.depth_list = sorted(depth_list)
        self.pixelshuffle_depth_list = sorted(pixelshuffle_depth_list)
        self.static_ks = 5 if kind == 's4' else 3
        D, S = max(self.depth_list), max(self.pixelshuffle_depth_list)
        groups, idx = [], 0
        self.mb_blocks = []
        if kind == 'x4':
            groups.append([0, 1])
            idx = 2
        n_mb_groups = 4 if kind == 's4' else 8
        for _ in range(n_mb_groups):
            groups.append(list(range(idx, idx + D)))
            self.mb_blocks += list(range(idx, idx + D))
            idx += D
        groups.append(list(range(idx, idx + S)))
        idx += S
        self.block_group_info = groups
        self.n_blocks = idx
        self.runtime_depth = [len(g) for g in groups]
        self.active_ks = {b: max(self.ks_list) for b in self.mb_blocks}
        self.active_e = {b: max(self.expand_ratio_list) for b in self.mb_blocks}

    # -- set / sample (Q2, Q3, Q4, Q6) ----------------------------------------------------------
    def set_active_subnet(self, ks=None, e=None, d=None, pixel_d=None):
        n_static = 2 if self.kind == 's4' else 4
        n_shuf_groups = 1 if self.kind == 's4' else 2
        as_list = lambda v, n: v if isinstance(v, list) else [v] * n
        ks = as_list(ks, self.n_blocks - n_static)
        e = as_list(e, self.n_blocks - n_static)
        depth = as_list(d, len(self.block_group_info) - n_shuf_groups)
        pix = as_list(pixel_d, n_shuf_groups)
        if self.kind == 's4':
            depth.insert(-1, pix[0])
            targets = list(range(0, self.n_blocks - 1))          # blocks[:-1]
        else:
            depth.insert(0, pix[0])
            depth.insert(-1, pix[0])
            targets = list(range(2, self.n_blocks - 2))          # blocks[2:-2]
        for b, k, ex in zip(targets, ks, e):
            if k is not None:
                self.active_ks[b] = k
            if ex is not None:
                self.active_e[b] = ex
        for i, dd in enumerate(depth):
            if dd is not None:
                self.runtime_depth[i] = min(len(self.block_group_info[i]), dd)

    def sample_active_subnet(self):
        n_static = 2 if self.kind == 's4' else 4
        n_shuf_groups = 1 if self.kind == 's4' else 2
        n = self.n_blocks - n_static
        ks = [random.choice(self.ks_list) for _ in range(n)]
        e = [random.choice(self.expand_ratio_list) for _ in range(n)]
        d = [random.choice(self.depth_list) for _ in range(len(self.block_group_info) - n_shuf_groups)]
        p = [random.choice(self.pixelshuffle_depth_list)]
        self.set_active_subnet(ks, e, d, p)
        return {'wid': None, 'ks': ks, 'e': e, 'd': d, 'pixel_d': p}

    # -- parameter inventory ---------------------------------------------------------------------
    def param_shapes(self):
        """state_dict keys -> shapes, in the reference's registration order."""
        k, kmax, emax = self.static_ks, max(self.ks_list), max(self.expand_ratio_list)
        mid = round(64 * emax)
        out = {}

        def bn(prefix, c):
            out[prefix + 'weight'] = (c,)
            out[prefix + 'bias'] = (c,)
            out[prefix + 'running_mean'] = (c,)
            out[prefix + 'running_var'] = (c,)
            out[prefix + 'num_batches_tracked'] = ()

        def conv_layer_(prefix, cin, cout):
            out[prefix + 'conv.weight'] = (cout, cin, k, k)
            bn(prefix + 'bn.', cout)

        def mb(prefix):
            p = prefix + 'mobile_inverted_conv.'
            out[p + 'inverted_bottleneck.conv.conv.weight'] = (mid, 64, 1, 1)
            bn(p + 'inverted_bottleneck.bn.bn.', mid)
            # a module's own parameters precede its children's in state_dict order
            for small, large in zip(self.ks_list[:-1], self.ks_list[1:]):
                out[p + 'depth_conv.conv.%dto%d_matrix' % (large, small)] = (small * small, small * small)
            out[p + 'depth_conv.conv.conv.weight'] = (mid, 1, kmax, kmax)
            bn(p + 'depth_conv.bn.bn.', mid)
            out[p + 'point_linear.conv.conv.weight'] = (64, mid, 1, 1)
            bn(p + 'point_linear.bn.bn.', 64)

        for b in range(self.n_blocks):
            if b in self.mb_blocks:
                mb('blocks.%d.' % b)
            elif self.kind == 'x4' and b == 0:
                conv_layer_('blocks.0.', 3, 16)
            elif self.kind == 'x4' and b == 1:
                conv_layer_('blocks.1.', 64, 16)
            else:
                conv_layer_('blocks.%d.' % b, 64, 256)
        if self.kind == 'x4':
            conv_layer_('enc_final_conv_blocks.0.', 64, 64)
            conv_layer_('enc_final_conv_blocks.1.', 64, 64)
            conv_layer_('enc_final_conv_blocks.2.', 64, 3)
        conv_layer_('dec_first_conv_block.', 3, 64)
        conv_layer_('dec_final_conv_blocks.0.', 64, 64)
        conv_layer_('dec_final_conv_blocks.1.', 64, 64)
        conv_layer_('dec_final_output_conv_block.', 64, 3)
        return out


def synth_state_dict(shapes, seed):
    """Deterministic, machine-independent synthetic parameters (numpy RandomState), following the
    SURVEY §8d recipe: he_fout conv weights, BN gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1),
    var~U(.5,1.5), transform matrices eye + 0.05 N(0,1) — so BN folding and the kernel transform are
    not vacuous."""
    rs = np.random.RandomState(seed)
    sd = {}
    for key, shape in shapes.items():
        if key.endswith('num_batches_tracked'):
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif key.endswith('_matrix'):
            n = shape[0]
            sd[key] = torch.from_numpy((np.eye(n) + 0.05 * rs.randn(n, n)).astype(np.float32))
        elif key.endswith('conv.weight'):
            fan = shape[0] * shape[2] * shape[3]
            sd[key] = torch.from_numpy((rs.randn(*shape) * math.sqrt(2.0 / fan)).astype(np.float32))
        elif key.endswith('running_var') or key.endswith('bn.weight'):
            sd[key] = torch.from_numpy(rs.uniform(0.5, 1.5, size=shape).astype(np.float32))
        elif key.endswith('running_mean') or key.endswith('bn.bias'):
            sd[key] = torch.from_numpy((0.1 * rs.randn(*shape)).astype(np.float32))
        else:
            raise KeyError(key)
    return sd


def _run_groups(x, sd, spec, lo, hi, training, transform_on):
    """Q1: the i-th group of the SLICE uses runtime_depth[i] (ofa_mbs4.py:149-153,162-166)."""
    groups = spec.block_group_info[lo:hi]
    for pos, block_idx in enumerate(groups):
        depth = spec.runtime_depth[pos]
        for b in block_idx[:depth]:
            prefix = 'blocks.%d.' % b
            if b in spec.mb_blocks:
                x = mbconv_block(x, sd, prefix, spec.active_ks[b], spec.active_e[b], spec.ks_list, training,
                                 transform_on=transform_on)
            else:
                act = 'pixelunshuffle' if (spec.kind == 'x4' and b < 2) else 'pixelshuffle'
                x = conv_layer(x, sd, prefix, act, training)
    return x


def supernet_forward(x, sd, spec, training=False, transform_on=True):
    """OFAMobileNetS4.forward (ofa_mbs4.py:142-178) / OFAMobileNetX4.forward (ofa_mbx4.py:185-254)."""
    t = training
    n_groups = len(spec.block_group_info)
    if spec.kind == 's4':
        x = conv_layer(x, sd, 'dec_first_conv_block.', None, t)
        skip = x
        x = _run_groups(x, sd, spec, 0, 4, t, transform_on)
        x = conv_layer(x, sd, 'dec_final_conv_blocks.0.', None, t) + skip
        x = conv_layer(x, sd, 'dec_final_conv_blocks.1.', None, t)
        x = _run_groups(x, sd, spec, 4, n_groups, t, transform_on)
        return conv_layer(x, sd, 'dec_final_output_conv_block.', None, t)
    x = _run_groups(x, sd, spec, 0, 1, t, transform_on)
    skip = x
    x = _run_groups(x, sd, spec, 1, 5, t, transform_on)
    x = conv_layer(x, sd, 'enc_final_conv_blocks.0.', None, t) + skip
    x = conv_layer(x, sd, 'enc_final_conv_blocks.1.', None, t)
    x = conv_layer(x, sd, 'enc_final_conv_blocks.2.', None, t)
    x = conv_layer(x, sd, 'dec_first_conv_block.', None, t)
    skip = x
    x = _run_groups(x, sd, spec, 5, 9, t, transform_on)
    x = conv_layer(x, sd, 'dec_final_conv_blocks.0.', None, t) + skip
    x = conv_layer(x, sd, 'dec_final_conv_blocks.1.', None, t)
    x = _run_groups(x, sd, spec, 9, n_groups, t, transform_on)
    return conv_layer(x, sd, 'dec_final_output_conv_block.', None, t)


# ---------------------------------------------------------------------------------------------
# SR quality metric (sr_run_manager.py:567-597, ofa/utils.py:27-34)
# ---------------------------------------------------------------------------------------------

def tensor_to_y_uint8(t):
    """[3,H,W] float in [0,1] -> uint8 BT.601 luma of the uint8-rounded RGB image."""
    img = (t.detach().float().cpu().clamp(0, 1).numpy().transpose(1, 2, 0) * 255.0).round().astype(np.uint8)
    y = (np.dot(img[..., :3], [65.481, 128.553, 24.966]) / 255.0 + 16.0).round()
    return y.astype(np.uint8)


def psnr_uint8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    if mse == 0:
        return float('inf')
    return 20 * math.log10(255.0 / math.sqrt(mse))


def y_uint8_batch(t):
    """rgb2y(tensor2img_np(.)) per image of a [N,3,H,W] batch (sr_run_manager.py:567-597): clamp to [0,1],
    x255 in fp32, round half to even, uint8; Y = round((65.481 R + 128.553 G + 24.966 B)/255 + 16) in float64."""
    t = t.detach().float().cpu().clamp(0, 1).numpy()
    img = (t * 255.0).round().astype(np.uint8).astype(np.float64)
    y = ((65.481 * img[:, 0] + 128.553 * img[:, 1] + 24.966 * img[:, 2]) / 255.0 + 16.0).round()
    return y.astype(np.uint8)


def psnr_y_sse(a, b):
    """Per-image sum of squared Y-uint8 differences (exact integers) of two [N,3,H,W] batches."""
    d = y_uint8_batch(a).astype(np.int64) - y_uint8_batch(b).astype(np.int64)
    return (d * d).reshape(d.shape[0], -1).sum(axis=1)


def psnr_y_from_sse(sse, n, h, w):
    """The reference's validate metric psnr(rgb2y(tensor2img_np(out)), rgb2y(tensor2img_np(hr)))
    (sr_run_manager.py:364,496; ofa/utils.py:27-34) from the per-image SSEs.  QUIRK kept: for a batch the
    reference first tiles the images with torchvision.utils.make_grid(nrow=int(sqrt(N)), padding=2), so the mean
    runs over the grid INCLUDING its zero padding (identical in both images: it adds pixels, not error)."""
    if n == 1:
        count = h * w
    else:
        nrow = int(math.sqrt(n))
        xmaps = min(nrow, n)
        ymaps = int(math.ceil(float(n) / xmaps))
        count = ((h + 2) * ymaps + 2) * ((w + 2) * xmaps + 2)
    mse = float(np.sum(sse)) / count
    if mse == 0:
        return float('inf')
    return 20 * math.log10(255.0 / math.sqrt(mse))


def set_running_statistics(sd, spec, batches):
    """elastic_nn/utils.py:16-66 — BatchNorm re-calibration of the active sub-network: forward every batch with
    each BN normalising by the batch's own statistics, then write the batch-size-weighted averages of the recorded
    means / biased variances into the first C entries of running_mean / running_var (in place in `sd`)."""
    global _RECAL_SINK
    _RECAL_SINK = []
    try:
        with torch.no_grad():
            for x in batches:
                supernet_forward(x, sd, spec)
        rec = _RECAL_SINK
    finally:
        _RECAL_SINK = None
    acc = {}
    for prefix, mean, var, n in rec:
        a = acc.setdefault(prefix, [0.0, 0.0, 0])
        a[0] = a[0] + mean * n
        a[1] = a[1] + var * n
        a[2] += n
    for prefix, (sm, sv, cnt) in acc.items():
        C = sm.shape[0]
        sd[prefix + 'running_mean'][:C] = sm / cnt
        sd[prefix + 'running_var'][:C] = sv / cnt
    return sd
