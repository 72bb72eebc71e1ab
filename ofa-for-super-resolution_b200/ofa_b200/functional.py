"""Operator layer between the drop-in nn.Modules and the C-ABI library.

Two families:
  * autograd Functions (training / grad-enabled path): DwConvFn, ConvFn, BnActFn, ReorderFn — forward
    and backward each call the library's kernels;
  * fused inference calls (no grad, module.eval()): conv + folded BN + activation + residual /
    long-skip + PixelShuffle / PixelUnshuffle store in ONE kernel, and the whole MBConv block.

Tensors produced here are logical NCHW stored channels-last (NHWC memory); inputs may have any
strides (the user's NCHW fp32 image is read in place).
"""
import ctypes
import os
from ctypes import byref

import torch

from . import backend as B

_state = {'compute_dtype': torch.float16, 'impl': B.IMPL_AUTO, 'mid_dtype': B.OFA_F16, 'train_dtype': torch.float32,
          'overflow': 'none', 'block_train': os.environ.get('OFA_BLOCK_TRAIN', '1') != '0',
          'train_side': 0 if os.environ.get('OFA_TRAIN_SIDE', '1') == '0' else 1}


def check_finite(t):
    """True when every element of `t` is finite (one device reduction + one 1-byte read)."""
    return bool(torch.isfinite(t).all().item())


def set_overflow_policy(policy):
    """fp16 activation storage (the default 16-bit format: it meets the 0.01 dB criterion) has a range of
    65504; the 64-channel trunk is an UNCLAMPED residual stream, so a checkpoint with very large BatchNorm gammas can
    leave it.  'none' (default): no check.  'fallback_bf16': a network-level inference forward whose fp16 result
    contains inf / NaN is recomputed with bf16 storage (fp32 range); costs one device -> host flag read per forward."""
    assert policy in ('none', 'fallback_bf16')
    _state['overflow'] = policy


def set_compute_dtype(dtype):
    """Activation storage type of the fused inference path: torch.float16 (tensor-core path, default:
    meets the 0.01 dB PSNR criterion; |activation| must stay below 65504), torch.bfloat16 (same
    kernels and speed, 3 fewer mantissa bits, fp32 range) or torch.float32 (exact CUDA-core path).
    Accumulation is fp32 in every mode."""
    assert dtype in (torch.bfloat16, torch.float16, torch.float32)
    _state['compute_dtype'] = dtype


def set_mid_dtype(dtype):
    """Storage type of the two expanded (ReLU6-clamped) intermediates of the planar MBConv path:
    torch.float16 (default: 3 more mantissa bits than bf16 on [0, 6]) or torch.bfloat16."""
    assert dtype in (torch.float16, torch.bfloat16)
    _state['mid_dtype'] = B.OFA_F16 if dtype == torch.float16 else B.OFA_BF16


def set_train_dtype(dtype):
    """Activation / activation-gradient storage of the autograd (training) path.  torch.float32 (default):
    the exact CUDA-core kernels.  torch.bfloat16: mixed precision — fp32 master weights, bf16 activations and
    gradients, forward convs and data gradients on the tcgen05 implicit-GEMM kernel, fp32 accumulation,
    fp32 weight gradients (progressive-shrinking training as BASELINE.json configs[2] names it)."""
    assert dtype in (torch.float32, torch.bfloat16)
    _state['train_dtype'] = dtype


def get_train_dtype():
    return _state['train_dtype']


def get_compute_dtype():
    return _state['compute_dtype']


def set_train_side_stream(on):
    """Weight gradients of the training MBConv blocks on a second, lower-priority CUDA stream beside the data-gradient chain
    (ofa_train_side_mode; fork / join inside every block call).  Default on; OFA_TRAIN_SIDE=0 or False here disables."""
    _state['train_side'] = 1 if on else 0
    if B._lib is not None:
        _side_mode_sync()


def _side_mode_sync():
    """Push the Python-side mode into the library (once the library is loaded)."""
    if _state.get('train_side_pushed') != _state['train_side']:
        B.lib().ofa_train_side_mode(_state['train_side'])
        _state['train_side_pushed'] = _state['train_side']


def set_block_train(on):
    """Training MBConv blocks as one library call each way (MBConvTrainFn, default on; OFA_BLOCK_TRAIN=0 or False here
    selects the layer-by-layer autograd nodes -- same kernels, same results, more host time)."""
    _state['block_train'] = bool(on)


def set_impl(impl):
    """Testing / profiling: force B.IMPL_SIMT or B.IMPL_FAST (default B.IMPL_AUTO)."""
    _state['impl'] = impl


def _stream(t):
    return B.stream_ptr(t.device)


# ---- optional per-call CUDA-event profiler (bench.py's live roofline measurement) -----------------
_profiler = None


def set_profiler(records):
    """`records` = a list that receives (tag, flops, algorithmic_bytes, start_event, end_event) for
    every fused inference call, or None to switch profiling off.  Events are recorded on the current
    stream, which is the stream the library launches on."""
    global _profiler
    _profiler = records


def _call(tag, flops, nbytes, fn):
    if _profiler is None:
        return fn()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    _profiler.append((tag, flops, nbytes, e0, e1))
    return out


def _null_or(t):
    return B.fptr(t) if t is not None else None


# =================================================================================================
# depthwise (a1, a2, a14)
# =================================================================================================

def _transform_ptrs(m75, m53):
    return _null_or(m75), _null_or(m53)


def dw_active_filter(w7, m75, m53, transform_on, ks, C):
    """DynamicSeparableConv2d.get_active_filter (dynamic_op.py:46-71) -> [C,1,ks,ks] fp32."""
    kmax = w7.shape[-1]
    out = torch.empty((C, 1, ks, ks), dtype=torch.float32, device=w7.device)
    p75, p53 = _transform_ptrs(m75, m53)
    B.check(B.lib().ofa_dw_active_filter(B.fptr(w7), kmax, p75, p53, int(bool(transform_on)), ks, C,
                                         out.data_ptr(), _stream(w7)))
    return out


def _dw_fwd_impl(x, w7, m75, m53, ks, transform_on):
    n, c, h, w = x.shape
    y = B.new_nhwc(n, c, h, w, x.dtype, x.device)
    p75, p53 = _transform_ptrs(m75, m53)
    tx, ty = B.t4(x), B.t4(y)
    B.check(B.lib().ofa_dw_fwd(byref(tx), byref(ty), B.fptr(w7), w7.shape[-1], p75, p53,
                               int(bool(transform_on)), ks, None, _state['impl'], _stream(x)))
    return y


def _dw_bwd_impl(x, w7, m75, m53, ks, transform_on, dy, need_dx, need_filter):
    """(dx, dw7, dm75, dm53) of the stride-1 elastic depthwise conv."""
    n, c, h, w = x.shape
    kmax = w7.shape[-1]
    p75, p53 = _transform_ptrs(m75, m53)
    L = B.lib()
    st = _stream(x)
    tdy = B.t4(dy)
    dx = dw7 = dm75 = dm53 = None
    if need_dx:
        dx = B.new_nhwc(n, c, h, w, x.dtype, dy.device)
        tdx = B.t4(dx)
        B.check(L.ofa_dw_bwd_data(byref(tdy), byref(tdx), B.fptr(w7), kmax, p75, p53,
                                  int(bool(transform_on)), ks, st))
    if need_filter:
        dwa = torch.empty((c, ks * ks), dtype=torch.float32, device=x.device)
        tx = B.t4(x)
        B.check(L.ofa_dw_bwd_filter(byref(tx), byref(tdy), ks, dwa.data_ptr(), st))
        dw7, dm75, dm53 = _dw_filter_chain(dwa, w7, m75, m53, transform_on, ks, c, st)
    return dx, dw7, dm75, dm53


class DwConvFn(torch.autograd.Function):
    """y = depthwise_conv(x, active_filter(w7, m75, m53, ks)), stride 1, same padding."""

    @staticmethod
    def forward(ctx, x, w7, m75, m53, ks, transform_on):
        y = _dw_fwd_impl(x, w7, m75, m53, ks, transform_on)
        ctx.save_for_backward(x, w7, m75, m53)
        ctx.ks, ctx.transform_on = ks, transform_on
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w7, m75, m53 = ctx.saved_tensors
        dx, dw7, dm75, dm53 = _dw_bwd_impl(x, w7, m75, m53, ctx.ks, ctx.transform_on, dy, ctx.needs_input_grad[0],
                                           ctx.needs_input_grad[1] or ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        return dx, dw7, dm75, dm53, None, None


def _dw_filter_chain(dwa, w7, m75, m53, transform_on, ks, c, st):
    """Gradient of the [C, ks*ks] active filter back to the stored 7x7 weights and the learned transform matrices
    (chain rule through dynamic_op.py:46-71)."""
    kmax = w7.shape[-1]
    p75, p53 = _transform_ptrs(m75, m53)
    dw7 = torch.zeros_like(w7)
    use75 = transform_on and ks < kmax and m75 is not None
    use53 = transform_on and ks < kmax and ks == 3 and m53 is not None
    dm75 = torch.zeros_like(m75) if m75 is not None else None
    dm53 = torch.zeros_like(m53) if m53 is not None else None
    B.check(B.lib().ofa_dw_active_filter_bwd(B.fptr(w7), kmax, p75, p53, int(bool(transform_on)), ks, c,
                                             dwa.data_ptr(), dw7.data_ptr(),
                                             dm75.data_ptr() if dm75 is not None else None,
                                             dm53.data_ptr() if dm53 is not None else None, st))
    return dw7, (dm75 if use75 else None), (dm53 if use53 else None)


class DwStridedConvFn(torch.autograd.Function):
    """Elastic depthwise conv with stride > 1 (DynamicSeparableConv2d built with stride 2 in the MobileNetV3-style
    nets, dynamic_op.py:73-84): same padding ks // 2, output (H - 1) // stride + 1.  Exact fp32-accumulate CUDA-core
    kernels on the materialised active filter."""

    @staticmethod
    def forward(ctx, x, w7, m75, m53, ks, transform_on, stride):
        n, c, h, w = x.shape
        filt = dw_active_filter(w7.detach(), m75, m53, transform_on, ks, c)
        y = B.new_nhwc(n, c, (h - 1) // stride + 1, (w - 1) // stride + 1, x.dtype, x.device)
        tx, ty = B.t4(x), B.t4(y)
        B.check(B.lib().ofa_dw_strided_fwd(byref(tx), byref(ty), filt.data_ptr(), ks, stride, None, _stream(x)))
        ctx.save_for_backward(x, w7, m75, m53, filt)
        ctx.cfg = (ks, transform_on, stride)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w7, m75, m53, filt = ctx.saved_tensors
        ks, transform_on, stride = ctx.cfg
        n, c, h, w = x.shape
        L, st = B.lib(), _stream(x)
        tdy = B.t4(dy)
        dx = dw7 = dm75 = dm53 = None
        if ctx.needs_input_grad[0]:
            dx = B.new_nhwc(n, c, h, w, x.dtype, dy.device)
            tdx = B.t4(dx)
            B.check(L.ofa_dw_strided_bwd_data(byref(tdy), byref(tdx), filt.data_ptr(), ks, stride, st))
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            dwa = torch.empty((c, ks * ks), dtype=torch.float32, device=x.device)
            tx = B.t4(x)
            B.check(L.ofa_dw_strided_bwd_filter(byref(tx), byref(tdy), ks, stride, dwa.data_ptr(), st))
            dw7, dm75, dm53 = _dw_filter_chain(dwa, w7, m75, m53, transform_on, ks, c, st)
        return dx, dw7, dm75, dm53, None, None, None


def dw_conv(x, w7, m75, m53, ks, transform_on, stride=1):
    if stride != 1:
        return DwStridedConvFn.apply(x, w7, m75, m53, ks, transform_on, stride)
    return DwConvFn.apply(x, w7, m75, m53, ks, transform_on)


# =================================================================================================
# sliced fully connected layer and squeeze-and-excite
# =================================================================================================

def _linear_fwd(x, w, ldw, bias, in_f, out_f, act):
    n = x.shape[0]
    y = torch.empty((n, out_f), dtype=torch.float32, device=x.device)
    B.check(B.lib().ofa_linear_fwd(x.data_ptr(), x.stride(0), B.fptr(w), ldw, _null_or(bias), n, in_f, out_f, act,
                                   y.data_ptr(), out_f, _stream(x)))
    return y


def _linear_bwd(dz, x, w, ldw, in_f, out_f, want_dx, has_bias):
    """(dx, dw_full, db_full): gradients of y = x w[:out,:in]^T + b[:out]; dw / db are full-size, zero outside the slice."""
    L, st = B.lib(), _stream(dz)
    n = dz.shape[0]
    dx = None
    if want_dx:
        dx = torch.empty((n, in_f), dtype=torch.float32, device=dz.device)
        B.check(L.ofa_linear_bwd_data(dz.data_ptr(), dz.stride(0), B.fptr(w), ldw, n, in_f, out_f, dx.data_ptr(), in_f, st))
    dw = torch.zeros_like(w)
    db = torch.zeros(w.shape[0], dtype=torch.float32, device=w.device) if has_bias else None
    B.check(L.ofa_linear_bwd_weight(dz.data_ptr(), dz.stride(0), x.data_ptr(), x.stride(0), n, in_f, out_f,
                                    dw.data_ptr(), ldw, db.data_ptr() if has_bias else None, st))
    return dx, dw, db


def _as_rows(x):
    if not x.is_cuda:
        raise RuntimeError('libofa_sr_b200 has no CPU path: tensor is on %s' % x.device)
    assert x.dim() == 2
    if x.dtype != torch.float32 or x.stride(1) != 1:
        x = x.float().contiguous()
    return x


class LinearFn(torch.autograd.Function):
    """DynamicLinear.forward (dynamic_op.py:115-136): F.linear(x, W[:out, :in], b[:out]) with the slice addressed in place."""

    @staticmethod
    def forward(ctx, x, w, bias, out_f):
        x = _as_rows(x)
        in_f = x.shape[1]
        ctx.save_for_backward(x, w)
        ctx.cfg = (in_f, out_f, bias is not None)
        return _linear_fwd(x, w, w.stride(0), bias, in_f, out_f, B.ACT_NONE)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        in_f, out_f, has_bias = ctx.cfg
        dx, dw, db = _linear_bwd(_as_rows(dy), x, w, w.stride(0), in_f, out_f, ctx.needs_input_grad[0], has_bias)
        return dx, dw, db, None


def linear(x, w, bias, out_features):
    return LinearFn.apply(x, w, bias, out_features)


class SEFn(torch.autograd.Function):
    """DynamicSE.forward (dynamic_op.py:175-200): y = x * h_sigmoid(W_e[:C,:mid] relu(W_r[:mid,:C] mean_hw(x) + b_r) + b_e).
    The 1x1 convs act on the pooled [N, C] matrix, i.e. they are sliced linear layers on the conv weights viewed
    as [out_max, in_max] matrices."""

    @staticmethod
    def forward(ctx, x, w_r, b_r, w_e, b_e, mid):
        n, c, h, w = x.shape
        L, st = B.lib(), _stream(x)
        tx = B.t4(x)
        pooled = torch.empty((n, c), dtype=torch.float32, device=x.device)
        B.check(L.ofa_plane_mean(byref(tx), pooled.data_ptr(), st))
        h1 = _linear_fwd(pooled, w_r, w_r.stride(0), b_r, c, mid, B.ACT_RELU)
        gate = _linear_fwd(h1, w_e, w_e.stride(0), b_e, mid, c, B.ACT_HSIGMOID)
        y = B.new_nhwc(n, c, h, w, x.dtype, x.device)
        ty = B.t4(y)
        B.check(L.ofa_channel_scale(byref(tx), byref(ty), gate.data_ptr(), None, st))
        ctx.save_for_backward(x, w_r, w_e, pooled, h1, gate)
        ctx.cfg = (mid, b_r is not None, b_e is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_r, w_e, pooled, h1, gate = ctx.saved_tensors
        mid, has_br, has_be = ctx.cfg
        n, c, h, w = x.shape
        L, st = B.lib(), _stream(x)
        tx, tdy = B.t4(x), B.t4(dy)
        dgate = torch.empty((n, c), dtype=torch.float32, device=x.device)
        B.check(L.ofa_plane_dot(byref(tx), byref(tdy), dgate.data_ptr(), st))
        dz2 = torch.empty_like(dgate)
        B.check(L.ofa_act_bwd_from_output(dgate.data_ptr(), gate.data_ptr(), B.ACT_HSIGMOID, dz2.data_ptr(), n * c, st))
        dh1, dw_e, db_e = _linear_bwd(dz2, h1, w_e, w_e.stride(0), mid, c, True, has_be)
        dz1 = torch.empty_like(dh1)
        B.check(L.ofa_act_bwd_from_output(dh1.data_ptr(), h1.data_ptr(), B.ACT_RELU, dz1.data_ptr(), n * mid, st))
        dpool, dw_r, db_r = _linear_bwd(dz1, pooled, w_r, w_r.stride(0), c, mid, True, has_br)
        dx = None
        if ctx.needs_input_grad[0]:
            dpool.mul_(1.0 / (h * w))             # the mean's gradient, spread over the plane by the kernel below
            dx = B.new_nhwc(n, c, h, w, x.dtype, x.device)
            tdx = B.t4(dx)
            B.check(L.ofa_channel_scale(byref(tdy), byref(tdx), gate.data_ptr(), dpool.data_ptr(), st))
        return dx, dw_r, db_r, dw_e, db_e, None


def squeeze_excite(x, w_reduce, b_reduce, w_expand, b_expand, mid):
    return SEFn.apply(x, w_reduce, b_reduce, w_expand, b_expand, mid)


# =================================================================================================
# dense conv with an in-place weight slice (a3, a4, a9, a14)
# =================================================================================================

def _conv_args(x, y, w, cin, cout, ks, store=B.STORE_PLAIN, epi=None, w_bf16=None, cin_pad=0, cout_pad=0):
    a = B.OfaConvArgs()
    a.x, a.y = B.t4(x), B.t4(y)
    a.w = B.fptr(w) if w is not None else None
    if w is not None:
        so, si, sh, sw = w.stride()
        a.w_so, a.w_si, a.w_sh, a.w_sw = so, si, sh, sw
    a.w_bf16 = w_bf16.data_ptr() if w_bf16 is not None else None
    a.cin_pad, a.cout_pad = cin_pad, cout_pad
    a.cin, a.cout, a.ks, a.flip, a.store = cin, cout, ks, 0, store
    if epi is not None:
        a.epi = epi
    return a


def _conv_out(x, cout, store, dtype, nchw=False):
    n, _, h, w = x.shape
    if store == B.STORE_PIXELSHUFFLE2:
        shape = (n, cout // 4, 2 * h, 2 * w)
    elif store == B.STORE_PIXELUNSHUFFLE2:
        shape = (n, cout * 4, h // 2, w // 2)
    else:
        shape = (n, cout, h, w)
    if nchw:
        return torch.empty(shape, dtype=dtype, device=x.device)
    return torch.empty(shape, dtype=dtype, device=x.device, memory_format=torch.channels_last)


def _train_cache(w, kind):
    """The packed 16-bit copies of a weight for the training path live ON the parameter object (one PackedWeightCache
    per layout `kind`), so their lifetime is the parameter's: a global table keyed by data_ptr would hand a new
    network the previous network's weights whenever the allocator reuses an address."""
    packs = w.__dict__.get('_ofa_packs')
    if packs is None:
        packs = w.__dict__['_ofa_packs'] = {}
    c = packs.get(kind)
    if c is None:
        c = packs[kind] = PackedWeightCache()
    return c


def _is_half_nhwc(t):
    return t.dtype in (torch.bfloat16, torch.float16) and t.is_contiguous(memory_format=torch.channels_last)


def _conv_fwd_impl(x, w, cin, cout, ks):
    """Forward of ConvFn (shared with the fused conv + BN + act node)."""
    tdt = _state['train_dtype']
    ydt = tdt if tdt != torch.float32 else x.dtype
    if cout < 16:
        ydt = torch.float32   # thin tensors (X4's 3-channel learned LR image) stay fp32, as in inference
    y = _conv_out(x, cout, B.STORE_PLAIN, ydt)
    impl, w16, cin_pad, cout_pad = B.IMPL_SIMT, None, 0, 0
    if tdt != torch.float32:
        if _is_half_nhwc(x) and cin % 64 == 0:
            cache = _train_cache(w, ('f', cin, cout))
            w16 = cache.get(w, cin, cout, ks, B.STORE_PLAIN, x.dtype)
            cin_pad, cout_pad, impl = cache.cin_pad, cache.cout_pad, B.IMPL_AUTO
        elif cin <= 4 and cout == 64:
            impl = B.IMPL_AUTO        # stem kernel: fp32 image in, 16-bit NHWC out
    a = _conv_args(x, y, w, cin, cout, ks, B.STORE_PLAIN, None, w16, cin_pad, cout_pad)
    B.check(B.lib().ofa_conv_fwd(byref(a), impl, _stream(x)))
    return y


def _conv_bwd_impl(x, w, cin, cout, ks, dy, need_dx, need_dw):
    """Backward of ConvFn: (dx, dw), either may be None."""
    L = B.lib()
    st = _stream(x)
    so, si, sh, sw = w.stride()
    if dy.dtype != torch.float32 and not dy.is_contiguous(memory_format=torch.channels_last):
        dy = dy.contiguous(memory_format=torch.channels_last)
    tdy = B.t4(dy)
    dx = dw = None
    if need_dx:
        n, _, h, wd = x.shape
        dx = B.new_nhwc(n, cin, h, wd, x.dtype, dy.device)     # gradients carry their activation's type
        if _is_half_nhwc(dy) and dy.dtype == x.dtype and cout % 64 == 0 and _state['train_dtype'] != torch.float32:
            # dX = conv(dY, W^T rotated 180 deg): pack W[o, i, ks-1-ky, ks-1-kx] as a (cout -> cin) weight
            cache = _train_cache(w, ('b', cin, cout))
            w16 = cache.get(w, cin, cout, ks, B.STORE_PLAIN, dy.dtype, rotated=True, backward=True)
            a = _conv_args(dy, dx, None, cout, cin, ks, B.STORE_PLAIN, None, w16, cache.cin_pad, cache.cout_pad)
            B.check(L.ofa_conv_fwd(byref(a), B.IMPL_FAST, st))
        elif (_state['train_dtype'] != torch.float32 and cout <= 4 and cin == 64 and ks in (3, 5)
              and dx.dtype != torch.float32 and w.is_contiguous()):
            # thin output (64 -> 3 at the SR resolution): its data gradient is a 3 -> 64 conv of dY with the rotated,
            # transposed slice -- the stem kernel, reading the fp32 master through swapped / negative strides
            last = (ks - 1) * sh + (ks - 1) * sw
            a = _conv_args(dy, dx, None, cout, cin, ks, B.STORE_PLAIN)
            a.w = w.data_ptr() + 4 * last
            a.w_so, a.w_si, a.w_sh, a.w_sw = si, so, -sh, -sw
            B.check(L.ofa_conv_fwd(byref(a), B.IMPL_AUTO, st))
        else:
            tdx = B.t4(dx)
            B.check(L.ofa_conv_bwd_data(byref(tdy), byref(tdx), B.fptr(w), so, si, sh, sw, cin, cout, ks, st))
    if need_dw:
        dw = torch.zeros_like(w)
        tdt = _state['train_dtype']
        if tdt != torch.float32 and cin == 64 and cout < 8 and _is_half_nhwc(x):
            # thin output (64 -> 3): pad dY to 8 channels so the tcgen05 weight-gradient kernel applies
            dy8 = torch.empty((dy.shape[0], 8, dy.shape[2], dy.shape[3]), dtype=x.dtype, device=dy.device,
                              memory_format=torch.channels_last).zero_()
            dy8[:, :cout] = dy
            dw8 = torch.zeros((8, w.shape[1], ks, ks), dtype=torch.float32, device=w.device)
            t8 = B.t4(dy8)
            tx = B.t4(x)
            B.check(L.ofa_conv_bwd_weight(byref(tx), byref(t8), dw8.data_ptr(), *dw8.stride(), cin, 8, ks, st))
            dw[:cout] = dw8[:cout]
        elif tdt != torch.float32 and cout == 64 and cin < 8 and _is_half_nhwc(dy):
            # thin input (the stem, 3 -> 64): pad X to 8 channels
            x8 = torch.empty((x.shape[0], 8, x.shape[2], x.shape[3]), dtype=dy.dtype, device=dy.device,
                             memory_format=torch.channels_last).zero_()
            x8[:, :cin] = x
            dw8 = torch.zeros((w.shape[0], 8, ks, ks), dtype=torch.float32, device=w.device)
            t8 = B.t4(x8)
            B.check(L.ofa_conv_bwd_weight(byref(t8), byref(tdy), dw8.data_ptr(), *dw8.stride(), 8, cout, ks, st))
            dw[:, :cin] = dw8[:, :cin]
        else:
            tx = B.t4(x)
            B.check(L.ofa_conv_bwd_weight(byref(tx), byref(tdy), dw.data_ptr(), so, si, sh, sw, cin, cout, ks, st))
    
    return dx, dw


class ConvFn(torch.autograd.Function):
    """y = conv2d(x, w[:cout, :cin]), stride 1, same padding; the slice is addressed in place.
    fp32 activations: the exact CUDA-core kernels.  16-bit NHWC activations (set_train_dtype): forward and
    data gradient run on the tcgen05 implicit-GEMM kernel (the data gradient is the same conv with the
    weight slice transposed and rotated by 180 degrees, packed from the fp32 master with swapped /
    negative strides); the weight gradient accumulates in fp32 into the slice."""

    @staticmethod
    def forward(ctx, x, w, cin, cout, ks):
        y = _conv_fwd_impl(x, w, cin, cout, ks)
        ctx.save_for_backward(x, w)
        ctx.dims = (cin, cout, ks)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        cin, cout, ks = ctx.dims
        dx, dw = _conv_bwd_impl(x, w, cin, cout, ks, dy, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dx, dw, None, None, None


def conv2d(x, w, cin, cout, ks):
    return ConvFn.apply(x, w, cin, cout, ks)


# =================================================================================================
# BatchNorm (+ activation, + residual) on the active channel prefix (a5, a6, a8, a14)
# =================================================================================================

def _bn_fwd_impl(x, gamma, beta, running_mean, running_var, residual, training, momentum, eps, act, bump):
    """Forward of BnActFn (shared with the fused nodes): returns (y, mean, var) -- the statistics used to normalise."""
    n, c, h, w = x.shape
    L = B.lib()
    st = _stream(x)
    tx = B.t4(x)
    if training and x.numel() > 0:
        # ONE library call: statistics, finalize (+ running-statistics slice update + num_batches_tracked bump in the same
        # launch), normalise + activation [+ residual]
        stats = torch.empty((2, c), dtype=torch.float32, device=x.device)
        mean, var = stats[0], stats[1]
        y = B.new_nhwc(n, c, h, w, x.dtype, x.device)
        ty = B.t4(y)
        update = running_mean is not None and momentum is not None and momentum != 0.0
        res_t4 = B.t4(residual) if residual is not None else None
        B.check(L.ofa_bn_train_fwd(byref(tx), byref(ty), _null_or(gamma), _null_or(beta),
                                   B.fptr(running_mean) if update else None, B.fptr(running_var) if update else None,
                                   float(momentum) if update else 0.0, float(eps), int(act),
                                   byref(res_t4) if res_t4 is not None else None, mean.data_ptr(), var.data_ptr(),
                                   bump.data_ptr() if (bump is not None and update) else None, st))
        if bump is not None and not update:
            bump += 1
        return y, mean, var
    mean, var = running_mean, running_var
    y = B.new_nhwc(n, c, h, w, x.dtype, x.device)
    ty = B.t4(y)
    e, keep = B.epilogue(gamma, beta, mean, var, eps, act, residual)
    B.check(L.ofa_affine_act(byref(tx), byref(ty), byref(e), B.STORE_PLAIN, st))
    return y, mean, var


def _bn_bwd_impl(x, gamma, beta, mean, var, training, eps, act, dy, need_dx, need_dgamma, need_dbeta):
    """Backward of BnActFn: (dx, dgamma, dbeta)."""
    n, c, h, w = x.shape
    L = B.lib()
    st = _stream(x)
    tx, tdy = B.t4(x), B.t4(dy)
    # the two per-channel sums ARE d(beta) and d(gamma) on the active prefix: they are reduced straight into one
    # zero-filled [2, C_full] buffer whose rows are returned as the full-width gradients (1 fill instead of 2 fills
    # + 2 slice copies per BatchNorm)
    c_full = max(c, gamma.shape[0] if gamma is not None else c, beta.shape[0] if beta is not None else c)
    sums = torch.zeros((2, c_full), dtype=torch.float32, device=x.device)
    s0, s1 = sums[0], sums[1]
    dx = dgamma = dbeta = None
    tdx = None
    if need_dx:
        dx = B.new_nhwc(n, c, h, w, x.dtype, dy.device)
        tdx = B.t4(dx)
    # reduce (+ apply) in one library call
    B.check(L.ofa_bn_train_bwd(byref(tx), byref(tdy), byref(tdx) if tdx is not None else None, _null_or(gamma),
                               _null_or(beta), B.fptr(mean), B.fptr(var), eps, act, int(training), s0.data_ptr(),
                               s1.data_ptr(), st))
    if gamma is not None and need_dgamma:
        dgamma = s1[:gamma.shape[0]]
    if beta is not None and need_dbeta:
        dbeta = s0[:beta.shape[0]]
    return dx, dgamma, dbeta


class BnActFn(torch.autograd.Function):
    """y = act(BN(x)) [+ residual].  Training: batch statistics (biased var for normalisation,
    unbiased for the running update, momentum update of the [:C] slice in place).  Eval: running
    statistics.  gamma / beta / running_* are the FULL-width tensors; the first C entries are used."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, training, momentum, eps, act, bump=None):
        y, mean, var = _bn_fwd_impl(x, gamma, beta, running_mean, running_var, residual, training, momentum, eps, act,
                                    bump)
        ctx.save_for_backward(x, gamma, beta, mean, var)
        ctx.cfg = (training, eps, act, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, var = ctx.saved_tensors
        training, eps, act, has_res = ctx.cfg
        dx, dgamma, dbeta = _bn_bwd_impl(x, gamma, beta, mean, var, training, eps, act, dy, ctx.needs_input_grad[0],
                                         ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        dres = dy if (has_res and ctx.needs_input_grad[5]) else None
        return dx, dgamma, dbeta, None, None, dres, None, None, None, None, None


class ConvBnActFn(torch.autograd.Function):
    """y = act(BN(conv2d(x, w[:cout, :cin]))) [+ residual] as ONE autograd node: the same library calls as ConvFn followed
    by BnActFn, but one Python-side node per layer instead of two (the eager training step is bounded by the host's
    launch rate; a node costs ~15 us each way)."""

    @staticmethod
    def forward(ctx, x, w, cin, cout, ks, gamma, beta, running_mean, running_var, residual, training, momentum, eps,
                act, bump):
        z = _conv_fwd_impl(x, w, cin, cout, ks)
        y, mean, var = _bn_fwd_impl(z, gamma, beta, running_mean, running_var, residual, training, momentum, eps, act,
                                    bump)
        ctx.save_for_backward(x, w, z, gamma, beta, mean, var)
        ctx.cfg = (cin, cout, ks, training, eps, act, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, z, gamma, beta, mean, var = ctx.saved_tensors
        cin, cout, ks, training, eps, act, has_res = ctx.cfg
        ng = ctx.needs_input_grad
        dz, dgamma, dbeta = _bn_bwd_impl(z, gamma, beta, mean, var, training, eps, act, dy, ng[0] or ng[1], ng[5], ng[6])
        dx = dw = None
        if dz is not None:
            dx, dw = _conv_bwd_impl(x, w, cin, cout, ks, dz, ng[0], ng[1])
        dres = dy if (has_res and ng[9]) else None
        return dx, dw, None, None, None, dgamma, dbeta, None, None, dres, None, None, None, None, None


class DwBnActFn(torch.autograd.Function):
    """y = act(BN(depthwise_conv(x, active_filter(w7, m75, m53, ks)))) as one autograd node (stride 1)."""

    @staticmethod
    def forward(ctx, x, w7, m75, m53, ks, transform_on, gamma, beta, running_mean, running_var, training, momentum, eps,
                act, bump):
        z = _dw_fwd_impl(x, w7, m75, m53, ks, transform_on)
        y, mean, var = _bn_fwd_impl(z, gamma, beta, running_mean, running_var, None, training, momentum, eps, act, bump)
        ctx.save_for_backward(x, w7, m75, m53, z, gamma, beta, mean, var)
        ctx.cfg = (ks, transform_on, training, eps, act)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w7, m75, m53, z, gamma, beta, mean, var = ctx.saved_tensors
        ks, transform_on, training, eps, act = ctx.cfg
        ng = ctx.needs_input_grad
        need_filter = ng[1] or ng[2] or ng[3]
        dz, dgamma, dbeta = _bn_bwd_impl(z, gamma, beta, mean, var, training, eps, act, dy, ng[0] or need_filter,
                                         ng[6], ng[7])
        dx = dw7 = dm75 = dm53 = None
        if dz is not None:
            dx, dw7, dm75, dm53 = _dw_bwd_impl(x, w7, m75, m53, ks, transform_on, dz, ng[0], need_filter)
        return dx, dw7, dm75, dm53, None, None, dgamma, dbeta, None, None, None, None, None, None, None


class MBConvTrainFn(torch.autograd.Function):
    """One MBConv block of the training step -- expand 1x1 -> BN -> act -> elastic depthwise -> BN -> act -> project 1x1
    -> BN [+ x] (dynamic_layers.py:70-84, proxyless_nets.py:44-51) -- as ONE autograd node and ONE library call each way
    (ofa_mbconv_train_fwd / _bwd: the same kernels in the same order as ConvBnActFn -> DwBnActFn -> ConvBnActFn).  The
    eager step was bounded by this interpreter loop, not by the GPU (tools/hosttime_train.py)."""

    @staticmethod
    def forward(ctx, x, w_exp, w7, m75, m53, w_proj, g1, b1, g2, b2, g3, b3, cfg):
        mid, cout, ks, transform_on, act, add_residual, bns = cfg
        n, cin, h, w = x.shape
        L = B.lib()
        _side_mode_sync()
        nbytes = L.ofa_mbconv_train_workspace_bytes(n, h, w, cin, mid, cout)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        y = B.new_nhwc(n, cout, h, w, x.dtype, x.device)
        a = B.OfaMBConvTrainArgs()
        a.x, a.y, a.dtype = x.data_ptr(), y.data_ptr(), B.dtype_code(x.dtype)
        a.n, a.h, a.w = n, h, w
        a.cin, a.mid, a.cout, a.ks, a.kmax, a.transform_on = cin, mid, cout, ks, w7.shape[-1], int(bool(transform_on))
        a.act, a.add_residual = int(act), int(bool(add_residual))
        a.w_exp = B.fptr(w_exp)
        a.w_exp_so, a.w_exp_si = w_exp.stride(0), w_exp.stride(1)
        a.w_dw, a.m75, a.m53 = B.fptr(w7), _null_or(m75), _null_or(m53)
        a.w_proj = B.fptr(w_proj)
        a.w_proj_so, a.w_proj_si = w_proj.stride(0), w_proj.stride(1)
        for slot, gamma, beta, (rm, rv, momentum, eps, bump) in ((a.bn_exp, g1, b1, bns[0]), (a.bn_dw, g2, b2, bns[1]),
                                                                  (a.bn_proj, g3, b3, bns[2])):
            update = rm is not None and momentum is not None and momentum != 0.0
            slot.gamma, slot.beta = B.fptr(gamma), B.fptr(beta)
            slot.running_mean = B.fptr(rm) if update else None
            slot.running_var = B.fptr(rv) if update else None
            slot.num_batches_tracked = bump.data_ptr() if (bump is not None and update) else None
            slot.momentum = float(momentum) if update else 0.0
            slot.eps = float(eps)
            if bump is not None and not update:
                bump += 1
        a.ws, a.ws_bytes = ws.data_ptr(), nbytes
        B.check(L.ofa_mbconv_train_fwd(byref(a), _stream(x)))
        ctx.save_for_backward(x, w_exp, w7, m75, m53, w_proj, g1, b1, g2, b2, g3, b3, ws)
        ctx.args = a
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_exp, w7, m75, m53, w_proj, g1, b1, g2, b2, g3, b3, ws = ctx.saved_tensors
        a = ctx.args
        if not dy.is_contiguous(memory_format=torch.channels_last) or dy.dtype != x.dtype:
            dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last)
        n, cin, h, w = x.shape
        dx = B.new_nhwc(n, cin, h, w, x.dtype, x.device)
        kmax = w7.shape[-1]
        use75 = bool(a.transform_on) and a.ks < kmax and m75 is not None
        use53 = bool(a.transform_on) and a.ks < kmax and a.ks == 3 and m53 is not None
        # every parameter gradient (full supernet size, zero outside the active slice) in ONE zero-filled buffer
        params = [w_exp, w7, w_proj, g1, b1, g2, b2, g3, b3]
        if m75 is not None:
            params.append(m75)
        if m53 is not None:
            params.append(m53)
        flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=x.device)
        parts = flat.split([p.numel() for p in params])
        base = flat.data_ptr()
        g = B.OfaMBConvTrainGrads()
        ptrs = []
        off = 0
        for p in params:
            ptrs.append(base + 4 * off)
            off += p.numel()
        g.dw_exp, g.dw_dw, g.dw_proj = ptrs[0], ptrs[1], ptrs[2]
        g.dgamma[0], g.dbeta[0], g.dgamma[1], g.dbeta[1], g.dgamma[2], g.dbeta[2] = ptrs[3:9]
        k = 9
        d75 = d53 = None
        if m75 is not None:
            g.dm75 = ptrs[k]
            d75 = parts[k].view(m75.shape) if use75 else None
            k += 1
        if m53 is not None:
            g.dm53 = ptrs[k]
            d53 = parts[k].view(m53.shape) if use53 else None
        B.check(B.lib().ofa_mbconv_train_bwd(byref(a), dy.data_ptr(), dx.data_ptr(), byref(g), _stream(x)))
        return (dx, parts[0].view(w_exp.shape), parts[1].view(w7.shape), d75, d53, parts[2].view(w_proj.shape),
                parts[3], parts[4], parts[5], parts[6], parts[7], parts[8], None)


def mbconv_train_supported(x, cin, mid, cout, residual, bns):
    """The block-level training call: 16-bit NHWC activations of the training dtype, 64-channel trunk, BatchNorms in
    batch-statistics mode without a set_running_statistics override."""
    tdt = _state['train_dtype']
    if tdt == torch.float32 or x.dtype != tdt or not _is_half_nhwc(x):
        return False
    if cin != 64 or cout != 64 or mid % 64 != 0 or not 64 <= mid <= 384:
        return False
    if residual is not None and residual is not x:
        return False
    if _profiler is not None or x.numel() == 0 or (x.data_ptr() & 15):
        return False
    for bn in bns:
        if bn is None or 'forward' in bn.__dict__ or bn.weight is None or bn.bias is None:
            return False
        if not (bn.training or not bn.track_running_stats):
            return False
    return True


def mbconv_train(x, w_exp, w7, m75, m53, w_proj, mid, cout, ks, transform_on, act, bn_exp, bn_dw, bn_proj, add_residual):
    modes = []
    for bn in (bn_exp, bn_dw, bn_proj):
        training, momentum, bump = _bn_mode(bn)
        modes.append((bn.running_mean, bn.running_var, momentum, bn.eps, bump))
    cfg = (mid, cout, ks, transform_on, act, add_residual, modes)
    return MBConvTrainFn.apply(x, w_exp, w7, m75, m53, w_proj, bn_exp.weight, bn_exp.bias, bn_dw.weight, bn_dw.bias,
                               bn_proj.weight, bn_proj.bias, cfg)


def bn_hooked(*bns):
    """True when any of the BatchNorm modules carries a per-instance forward override (set_running_statistics)."""
    return any(b is not None and 'forward' in b.__dict__ for b in bns)


def act_residual(y, act=B.ACT_NONE, residual=None):
    """y = act(y) + residual through the elementwise kernel (no BN)."""
    if act == B.ACT_NONE and residual is None:
        return y
    n, c, h, w = y.shape
    out = B.new_nhwc(n, c, h, w, y.dtype, y.device)
    if out.numel() == 0:
        return out
    e, keep = B.epilogue(act=act, residual=residual)
    ty, to = B.t4(y), B.t4(out)
    B.check(B.lib().ofa_affine_act(byref(ty), byref(to), byref(e), B.STORE_PLAIN, _stream(y)))
    return out


def batch_stats(x):
    """Per-channel batch mean and BIASED variance of x [N,C,H,W] on the device (ofa_bn_stats): fp32 [C] each."""
    c = x.shape[1]
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    var = torch.empty(c, dtype=torch.float32, device=x.device)
    tx = B.t4(x)
    B.check(B.lib().ofa_bn_stats(byref(tx), mean.data_ptr(), var.data_ptr(), _stream(x)))
    return mean, var


def normalize_with(x, mean, var, gamma, beta, eps):
    """F.batch_norm(x, mean, var, gamma[:C], beta[:C], False, 0.0, eps) through the elementwise kernel."""
    n, c, h, w = x.shape
    y = B.new_nhwc(n, c, h, w, x.dtype, x.device)
    e, keep = B.epilogue(gamma, beta, mean, var, eps, B.ACT_NONE, None)
    tx, ty = B.t4(x), B.t4(y)
    B.check(B.lib().ofa_affine_act(byref(tx), byref(ty), byref(e), B.STORE_PLAIN, _stream(x)))
    return y


def bn_act(x, bn, C, act=B.ACT_NONE, residual=None, full_width=False):
    """DynamicBatchNorm2d.bn_forward semantics (dynamic_op.py:148-167) for an nn.BatchNorm2d `bn`
    on the first C channels.  `full_width` = the reference's `bn(x)` branch, which also bumps
    num_batches_tracked through nn.BatchNorm2d.forward."""
    if 'forward' in bn.__dict__:
        # a per-instance `forward` override = the hook set_running_statistics installs on every BatchNorm2d of a
        # deep copy (reference elastic_nn/utils.py:29-52): honour it, then apply the rest of the fused epilogue
        return act_residual(bn(x), act, residual)
    training, momentum, bump = _bn_mode(bn)
    return BnActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual, training,
                         momentum, bn.eps, act, bump)


def _bn_mode(bn):
    """(training, momentum, bump) of nn.BatchNorm2d.forward / DynamicBatchNorm2d.bn_forward (dynamic_op.py:148-167)."""
    training = bn.training or not bn.track_running_stats
    momentum = 0.0
    bump = None
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        if bn.momentum is None:
            bn.num_batches_tracked += 1
            momentum = 1.0 / float(bn.num_batches_tracked)
        else:
            momentum = bn.momentum
            bump = bn.num_batches_tracked          # incremented by the running-statistics kernel (one launch less)
            if not (bump.is_cuda and bump.dtype == torch.int64):
                bn.num_batches_tracked += 1
                bump = None
    return training, momentum, bump


def conv_bn_act(x, w, cin, cout, ks, bn, act=B.ACT_NONE, residual=None):
    """conv2d -> bn_act as one autograd node (two when the BatchNorm carries a set_running_statistics override)."""
    if 'forward' in bn.__dict__:
        return bn_act(conv2d(x, w, cin, cout, ks), bn, cout, act, residual)
    training, momentum, bump = _bn_mode(bn)
    return ConvBnActFn.apply(x, w, cin, cout, ks, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual,
                             training, momentum, bn.eps, act, bump)


def dw_bn_act(x, w7, m75, m53, ks, transform_on, bn, act=B.ACT_NONE, stride=1):
    """depthwise conv -> bn_act as one autograd node (stride 1, no override on the BatchNorm), else two."""
    if stride != 1 or 'forward' in bn.__dict__:
        return bn_act(dw_conv(x, w7, m75, m53, ks, transform_on, stride), bn, x.shape[1], act)
    training, momentum, bump = _bn_mode(bn)
    return DwBnActFn.apply(x, w7, m75, m53, ks, transform_on, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                           training, momentum, bn.eps, act, bump)


# =================================================================================================
# PixelShuffle(2) / PixelUnshuffle(2) as a stand-alone reorder (training path; a10, a11)
# =================================================================================================

class ReorderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, store):
        n, c, h, w = x.shape
        y = _conv_out(x, c, store, x.dtype)
        tx, ty = B.t4(x), B.t4(y)
        B.check(B.lib().ofa_affine_act(byref(tx), byref(ty), None, store, _stream(x)))
        ctx.store = store
        return y

    @staticmethod
    def backward(ctx, dy):
        inv = B.STORE_PIXELUNSHUFFLE2 if ctx.store == B.STORE_PIXELSHUFFLE2 else B.STORE_PIXELSHUFFLE2
        n, c, h, w = dy.shape
        dx = _conv_out(dy, c, inv, dy.dtype)
        tdy, tdx = B.t4(dy), B.t4(dx)
        B.check(B.lib().ofa_affine_act(byref(tdy), byref(tdx), None, inv, _stream(dy)))
        return dx, None


def pixel_shuffle2(x):
    return ReorderFn.apply(x, B.STORE_PIXELSHUFFLE2)


def pixel_unshuffle2(x):
    return ReorderFn.apply(x, B.STORE_PIXELUNSHUFFLE2)


# =================================================================================================
# fused inference calls
# =================================================================================================

# ---- derived 16-bit weight copies ---------------------------------------------------------------------------------
# The tensor-core kernels read packed 16-bit copies of the active fp32 weight slices.  A copy is NEVER trusted across
# forward passes: `Tensor._version` (what round 1 keyed on) does not move for `w.data.copy_()` / `w.data.normal_()`
# (ofa/utils.py:134-155 init_model, elastic_nn/utils.py:76-82), raw-pointer writers or NCCL broadcasts, so any cache
# keyed on it can serve stale weights.  Instead every top-level forward pass (`forward_scope`) starts a new EPOCH on
# its device and a copy is fresh only if it was written in the current epoch:
#   * the first pass with a given (sub-network, shape, dtype) signature packs each copy on first use and records the
#     jobs as a PLAN; every later pass with that signature re-derives ALL of its copies from the fp32 masters with ONE
#     multi-job launch (ofa_pack_weights_multi: ~16 MB of traffic, a few microseconds) before the first layer runs --
#     inside a captured CUDA graph too, so graph replays follow weight updates;
#   * passes without a plan (training with a freshly sampled sub-network, stand-alone modules) pack on first use per
#     epoch, one small launch per copy -- what an optimizer step forced anyway;
#   * the backward pass belongs to the epoch of the last forward on that device.
# Slots are per device, so nn.DataParallel replicas (shallow module copies sharing this object, one thread per GPU,
# sr_run_manager.py:197-198) never hand each other buffers.
import threading

_tls = threading.local()
_epochs = {}                 # device index -> epoch counter
_generation = [0]            # bumped by invalidate_packed_weights()
_tables_keepalive = []       # device job tables are never freed: a captured CUDA graph may still replay them
_MAX_PLANS = 64


def invalidate_packed_weights():
    """Explicitly drop every derived 16-bit weight copy (they are re-derived on next use).  Not needed for correctness
    -- every forward pass re-derives its copies -- kept as the explicit hook init_model / load_state_dict /
    re_organize_middle_weights / broadcast_parameters call."""
    _generation[0] += 1


def _token(di):
    return (_generation[0], _epochs.get(di, 0))


def _new_epoch(di):
    _epochs[di] = _epochs.get(di, 0) + 1


class _Plan:
    __slots__ = ('jobs', 'table', 'n')

    def __init__(self, jobs, device):
        self.jobs = jobs                                        # [(cache, slot_key, w, wptr, buf, cin_pad, cout_pad)]
        arr = (B.OfaPackJob * len(jobs))()
        for k, (cache, key, w, job, buf, cp, op) in enumerate(jobs):
            arr[k] = job
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = raw.to(device)
        _tables_keepalive.append(self.table)
        self.n = len(jobs)

    def replay(self, di, stream):
        """Re-derive every copy of the plan (one launch); False when a master moved (plan must be re-recorded)."""
        for cache, key, w, job, buf, cp, op in self.jobs:
            if w.data_ptr() != key[0]:
                return False
        B.check(B.lib().ofa_pack_weights_multi(self.table.data_ptr(), self.n, stream))
        tok = _token(di)
        for cache, key, w, job, buf, cp, op in self.jobs:
            cache._slots[di] = [key, buf, cp, op, tok]
        return True


class forward_scope:
    """`with forward_scope(module, x[, signature]):` around a module's forward.  The OUTERMOST scope of a thread starts
    a new weight epoch on x's device; with a hashable `signature` (networks pass their active sub-network + input
    shape + dtypes) it replays / records the pack plan described above."""

    def __init__(self, owner, x, signature=None):
        self.owner, self.x, self.signature = owner, x, signature
        self.top = False
        self.rec = None

    def __enter__(self):
        depth = getattr(_tls, 'depth', 0)
        _tls.depth = depth + 1
        x = self.x
        if depth != 0 or not (torch.is_tensor(x) and x.is_cuda):
            return self
        self.top = True
        di = x.device.index
        _new_epoch(di)
        if self.signature is None:
            return self
        plans = self.owner.__dict__.setdefault('_ofa_pack_plans', {})
        pkey = (di, self.signature)
        plan = plans.get(pkey)
        if plan is not None:
            if plan.replay(di, _stream(x)):
                return self
            del plans[pkey]
        if len(plans) < _MAX_PLANS and not torch.cuda.is_current_stream_capturing():
            self.rec = _tls.rec = []
        return self

    def __exit__(self, et, ev, tb):
        _tls.depth -= 1
        if self.top and self.rec is not None:
            _tls.rec = None
            if et is None and self.rec:
                plans = self.owner.__dict__.setdefault('_ofa_pack_plans', {})
                plans[(self.x.device.index, self.signature)] = _Plan(self.rec, self.x.device)
        return False


def scoped_forward(signature=None):
    """Decorator for a module's `forward(self, x, ...)`: runs it inside a forward_scope (nested calls pay one
    attribute lookup).  `signature(self, x)` -> hashable plan key, or None for no plan."""
    def wrap(fwd):
        def forward(self, x, *args, **kwargs):
            if getattr(_tls, 'depth', 0):
                return fwd(self, x, *args, **kwargs)
            sig = signature(self, x) if signature is not None else None
            with forward_scope(self, x, sig):
                y = fwd(self, x, *args, **kwargs)
            if (signature is not None and _state['overflow'] == 'fallback_bf16' and _state['compute_dtype'] == torch.float16
                    and torch.is_tensor(y) and y.is_cuda and not torch.is_grad_enabled() and not self.training
                    and not torch.cuda.is_current_stream_capturing() and not check_finite(y)):
                _state['compute_dtype'] = torch.bfloat16
                try:
                    sig = signature(self, x)
                    with forward_scope(self, x, sig):
                        y = fwd(self, x, *args, **kwargs)
                finally:
                    _state['compute_dtype'] = torch.float16
            return y
        forward.__doc__ = fwd.__doc__
        forward.__wrapped__ = fwd
        return forward
    return wrap


def pack_plan_signature(net, x):
    """Everything that decides WHICH weight copies a network forward uses: the active sub-network, the input shape and
    the dtype / dispatch state (a wrong guess only costs speed: copies outside the plan are packed on first use)."""
    if not (torch.is_tensor(x) and x.is_cuda) or net.training or torch.is_grad_enabled():
        return None
    blocks = []
    for blk in net.blocks:
        mb = getattr(blk, 'mobile_inverted_conv', None)
        if mb is not None:
            blocks.append((getattr(mb, 'active_kernel_size', 0), getattr(mb, 'active_expand_ratio', 0)))
    return (tuple(x.shape), x.dtype, _state['compute_dtype'], _state['mid_dtype'], _state['impl'], _profiler is None,
            tuple(net.runtime_depth), tuple(blocks))


class PackedWeightCache:
    """16-bit [tap][cout_pad][cin_pad] copy (in the activation's format) of an active weight slice, owned by the module (or,
    for the training path, by the parameter).  See the epoch rules above: `get` re-derives the copy unless it was
    already written during the current forward pass."""

    def __init__(self):
        self._slots = {}            # device index -> [key, buf, cin_pad, cout_pad, token]
        self.cin_pad = self.cout_pad = 0

    def get(self, w, cin, cout, ks, store, dtype=torch.bfloat16, cout_pad=None, rotated=False, backward=False):
        """rotated=True packs the data-gradient weight: W[o, i, ks-1-ky, ks-1-kx] as a (cout -> cin) conv weight, read
        from the fp32 master through swapped / negative strides.  backward=True marks a call from an autograd backward
        (it belongs to the epoch of the forward pass that preceded it)."""
        di = w.device.index
        if getattr(_tls, 'depth', 0) == 0 and not backward:
            _new_epoch(di)          # a bare functional call is its own forward pass
        key = (w.data_ptr(), tuple(w.shape), w.stride(), cin, cout, ks, store, dtype, cout_pad, rotated)
        slot = self._slots.get(di)
        tok = _token(di)
        if slot is not None and slot[0] == key and slot[4] == tok:
            self.cin_pad, self.cout_pad = slot[2], slot[3]
            return slot[1]
        so, si, sh, sw = w.stride()
        if rotated:
            p_cin, p_cout = cout, cin
            cin_pad, c_pad = cout, (cin + 15) // 16 * 16
            wptr = w.data_ptr() + 4 * ((ks - 1) * sh + (ks - 1) * sw)
            strides = (si, so, -sh, -sw)
        else:
            p_cin, p_cout = cin, cout
            cin_pad = (cin + 63) // 64 * 64
            c_pad = cout_pad if cout_pad is not None else (cout + 15) // 16 * 16
            wptr = B.fptr(w)
            strides = (so, si, sh, sw)
        shape = (ks * ks, c_pad, cin_pad)
        buf = slot[1] if slot is not None else None
        if buf is None or tuple(buf.shape) != shape or buf.dtype != dtype:
            buf = torch.empty(shape, dtype=dtype, device=w.device)   # else: repack in place (stream-ordered)
        job = B.OfaPackJob(wptr, strides[0], strides[1], strides[2], strides[3], p_cin, p_cout, ks, cin_pad, c_pad,
                           store, B.dtype_code(dtype), 0, buf.data_ptr())
        B.check(B.lib().ofa_pack_weight_16(job.w, job.w_so, job.w_si, job.w_sh, job.w_sw, p_cin, p_cout, ks, cin_pad,
                                           c_pad, store, job.dtype, job.out, _stream(w)))
        self._slots[di] = [key, buf, cin_pad, c_pad, tok]
        self.cin_pad, self.cout_pad = cin_pad, c_pad
        rec = getattr(_tls, 'rec', None)
        if rec is not None and not backward:
            rec.append((self, key, w, job, buf, cin_pad, c_pad))
        return buf


def _bn_epilogue(bn, act, residual):
    if bn is None:
        return B.epilogue(act=act, residual=residual)
    return B.epilogue(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, act, residual)


def conv_bn_act_infer(x, w, cin, cout, ks, bn=None, act=B.ACT_NONE, store=B.STORE_PLAIN, residual=None,
                      cache=None, out_dtype=None, out_nchw=False):
    """ConvLayer / DynamicPointConv2d(+BN+act) in inference: one kernel."""
    dtype = out_dtype or _state['compute_dtype']
    y = _conv_out(x, cout, store, dtype, nchw=out_nchw)
    if y.numel() == 0:            # empty batch / empty image: nothing to launch (F.conv2d returns an empty tensor too)
        return y
    e, keep = _bn_epilogue(bn, act, residual)
    w_bf16, cin_pad, cout_pad = None, 0, 0
    impl = _state['impl']
    if x.dtype in (torch.bfloat16, torch.float16) and cin % 64 == 0 and cache is not None and impl != B.IMPL_SIMT:
        w_bf16 = cache.get(w, cin, cout, ks, store, x.dtype)
        cin_pad, cout_pad = cache.cin_pad, cache.cout_pad
    elif impl == B.IMPL_FAST:
        impl = B.IMPL_AUTO  # layers the tensor-core kernel does not cover (thin stem) use the CUDA-core one
    a = _conv_args(x, y, w, cin, cout, ks, store, e, w_bf16, cin_pad, cout_pad)
    P = x.shape[0] * x.shape[2] * x.shape[3]
    nbytes = P * (cin * x.element_size() + cout * y.element_size()) + \
        (y.numel() * residual.element_size() if residual is not None else 0)
    half_nhwc = x.dtype in (torch.bfloat16, torch.float16) and x.is_contiguous(memory_format=torch.channels_last)
    if impl != B.IMPL_SIMT and store == B.STORE_PLAIN and half_nhwc and cin == 64 and ks in (3, 5) and ks * cout <= 16:
        kind = 'rows-tc'          # conv_out_rows_kernel (kx folded into the accumulator columns)
    elif impl != B.IMPL_SIMT and cin <= 4 and ks in (3, 5) and ((cout == 64 and store == B.STORE_PLAIN) or
                                                                  (cout == 16 and store == B.STORE_PIXELUNSHUFFLE2)):
        kind = 'stem'             # conv_stem_kernel
    else:
        kind = 'tc' if w_bf16 is not None else 'simt'
    tag = 'conv%dx%d %d->%d%s %s' % (ks, ks, cin, cout, {0: '', 1: '+ps', 2: '+pus'}[store], kind)
    _call(tag, 2.0 * P * ks * ks * cin * cout, nbytes,
          lambda: B.check(B.lib().ofa_conv_fwd(byref(a), impl, _stream(x))))
    return y


def dw_bn_act_infer(x, w7, m75, m53, ks, transform_on, bn, act):
    n, c, h, w = x.shape
    y = B.new_nhwc(n, c, h, w, x.dtype, x.device)
    if y.numel() == 0:
        return y
    e, keep = _bn_epilogue(bn, act, None)
    p75, p53 = _transform_ptrs(m75, m53)
    tx, ty = B.t4(x), B.t4(y)
    _call('dw%dx%d C%d' % (ks, ks, c), 2.0 * n * h * w * c * ks * ks, 2 * n * h * w * c * x.element_size(),
          lambda: B.check(B.lib().ofa_dw_fwd(byref(tx), byref(ty), B.fptr(w7), w7.shape[-1], p75, p53,
                                             int(bool(transform_on)), ks, byref(e), _state['impl'], _stream(x))))
    return y




def _bn_struct(bn):
    return B.OfaBn(B.fptr(bn.weight), B.fptr(bn.bias), B.fptr(bn.running_mean), B.fptr(bn.running_var), bn.eps)


def mbconv_infer(x, w_exp, w_dw, m75, m53, w_proj, cin, mid, cout, ks, transform_on, act, bn_exp, bn_dw,
                 bn_proj, add_residual, pack_cache=None):
    """Whole inference MBConv block (expand -> dw -> project [+x]) through ofa_mbconv_fwd; needs
    NHWC-dense bf16 x.  Returns NHWC bf16."""
    n, _, h, w = x.shape
    if x.numel() == 0:
        return B.new_nhwc(n, cout, h, w, x.dtype, x.device)
    planar = (planar_supported(x, cin, mid, cout) and _state['impl'] not in (B.IMPL_SIMT, B.IMPL_NHWC)
              and (_state['impl'] in (B.IMPL_FAST, B.IMPL_BAND, B.IMPL_PLANAR3) or planar_preferred(x)))
    band = planar and act == B.ACT_RELU6 and _state['impl'] != B.IMPL_PLANAR3 and (
        _state['impl'] == B.IMPL_BAND or band_preferred(x))
    if _profiler is not None and planar and not band:
        return _mbconv_planar_staged(x, w_exp, w_dw, m75, m53, w_proj, mid, ks, transform_on, act, bn_exp, bn_dw,
                                     bn_proj, add_residual)
    if _profiler is not None and not planar:
        # same kernels, issued one by one so each gets its own event pair
        c1 = _train_cache(w_exp, ('prof', 0))
        c2 = _train_cache(w_proj, ('prof', 1))
        t = conv_bn_act_infer(x, w_exp, cin, mid, 1, bn_exp, act, cache=c1)
        t = dw_bn_act_infer(t, w_dw, m75, m53, ks, transform_on, bn_dw, act)
        return conv_bn_act_infer(t, w_proj, mid, cout, 1, bn_proj, B.ACT_NONE, residual=x if add_residual else None,
                                 cache=c2)
    y = B.new_nhwc(n, cout, h, w, x.dtype, x.device)
    L = B.lib()
    ws_bytes = L.ofa_mbconv_workspace_bytes(n, h, w, cin, mid, cout)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    a = B.OfaMBConvArgs()
    a.x, a.y = B.t4(x), B.t4(y)
    a.w_exp = B.fptr(w_exp)
    a.w_exp_so, a.w_exp_si = w_exp.stride(0), w_exp.stride(1)
    a.w_dw, a.kmax = B.fptr(w_dw), w_dw.shape[-1]
    a.m75, a.m53 = _transform_ptrs(m75, m53)
    a.transform_on = int(bool(transform_on))
    a.w_proj = B.fptr(w_proj)
    a.w_proj_so, a.w_proj_si = w_proj.stride(0), w_proj.stride(1)
    a.cin, a.mid, a.cout, a.ks, a.act = cin, mid, cout, ks, act
    a.bn_exp, a.bn_dw, a.bn_proj = _bn_struct(bn_exp), _bn_struct(bn_dw), _bn_struct(bn_proj)
    a.add_residual = int(bool(add_residual))
    a.ws, a.ws_bytes = ws.data_ptr(), ws_bytes
    a.mid_dtype = _state['mid_dtype']
    if planar and pack_cache is not None:
        we, wp = _planar_packed_weights(pack_cache, w_exp, w_proj, mid, x.dtype)
        a.w_exp_packed, a.w_proj_packed = we.data_ptr(), wp.data_ptr()
    if band:
        # one launch for the whole block: algorithmic traffic = trunk in + out (SURVEY 8d "fused MBConv block")
        P = n * h * w
        _call('mbconv band ks%d M%d' % (ks, mid), 2.0 * P * mid * (128 + ks * ks), 2 * P * 64 * x.element_size(),
              lambda: B.check(L.ofa_mbconv_fwd(byref(a), _state['impl'], _stream(x))))
        return y
    B.check(L.ofa_mbconv_fwd(byref(a), _state['impl'], _stream(x)))
    return y


def _planar_packed_weights(cache, w_exp, w_proj, mid, trunk_dtype):
    """Planar-format 16-bit copies of a block's (expand, project) weight slices: expand = [ceil128(mid)][64] in the trunk's
    format (the UMMA A operand, zero rows above `mid`), project = [64][mid] in the intermediates' format.  Both are plain
    ks = 1 packs, so they live in PackedWeightCache slots (`cache` is a dict owned by the module) and follow the same
    epoch / plan rules as every other derived weight copy."""
    mdt = _state['mid_dtype']
    tdt = torch.float16 if mdt in (0, B.OFA_F16) else torch.bfloat16
    ce = cache.get('exp')
    if ce is None:
        ce = cache.setdefault('exp', PackedWeightCache())
    cp = cache.get('proj')
    if cp is None:
        cp = cache.setdefault('proj', PackedWeightCache())
    we = ce.get(w_exp, 64, mid, 1, B.STORE_PLAIN, trunk_dtype, cout_pad=(mid + 127) // 128 * 128)
    wp = cp.get(w_proj, mid, 64, 1, B.STORE_PLAIN, tdt, cout_pad=64)
    return we, wp


def planar_supported(x, cin, mid, cout):
    """Shapes the planar tcgen05 MBConv path takes (mirrors mbconv_planar_supported in the library)."""
    return (x.dtype in (torch.bfloat16, torch.float16) and cin == 64 and cout == 64 and mid % 64 == 0 and 64 <= mid <= 384
            and x.shape[3] % 8 == 0)


def planar_preferred(x):
    """IMPL_AUTO's choice (mirrors mbconv_planar_preferred): the planar depthwise rebuilds its filter matrices per channel
    plane and works in 128 (64) x 112-pixel tiles, so planes under 2304 pixels (48 x 48) or filling < 25 % of their tiles -- batches
    of small patches -- take the three NHWC kernels; IMPL_FAST forces the planar path."""
    h, w = x.shape[2], x.shape[3]
    tail = h % 128
    rows = h // 128 * 128 + (0 if tail == 0 else 64 if tail <= 64 else 128)
    cols = (w + 111) // 112 * 112
    return h * w >= 2304 and 4 * h * w >= rows * cols


def band_preferred(x):
    """Mirrors mbconv_band_preferred: the single-launch, L2-resident block is opt-in (OFA_BAND_ENABLE=1 or IMPL_BAND): at
    the bench shape it measured slower than the three stand-alone kernels."""
    import os
    h, w = x.shape[2], x.shape[3]
    return (int(os.environ.get('OFA_BAND_ENABLE', '0') or 0) and planar_preferred(x) and h * w >= 128 * 448 and w >= 224
            and h >= 96 and not int(os.environ.get('OFA_BAND_DISABLE', '0') or 0))


def _mbconv_planar_staged(x, w_exp, w_dw, m75, m53, w_proj, mid, ks, transform_on, act, bn_exp, bn_dw, bn_proj,
                          add_residual):
    """The planar path stage by stage through the exported stage entry points (what ofa_mbconv_fwd runs
    internally), so that the profiler gets one event pair per kernel."""
    n, _, h, w = x.shape
    hw = h * w
    L = B.lib()
    st = _stream(x)
    dt = _state['mid_dtype']
    tdt = torch.float16 if dt == B.OFA_F16 else torch.bfloat16
    xdt = B.dtype_code(x.dtype)
    we = torch.empty(((mid + 127) // 128 * 128, 64), dtype=x.dtype, device=x.device)
    wp = torch.empty((64, mid), dtype=tdt, device=x.device)
    B.check(L.ofa_mbconv_pack_weights(B.fptr(w_exp), w_exp.stride(0), w_exp.stride(1), B.fptr(w_proj),
                                      w_proj.stride(0), w_proj.stride(1), mid, xdt, dt, we.data_ptr(), wp.data_ptr(),
                                      st))
    t1 = torch.empty((n, mid, hw), dtype=tdt, device=x.device)
    t2 = torch.empty((n, mid, hw), dtype=tdt, device=x.device)
    y = B.new_nhwc(n, 64, h, w, x.dtype, x.device)
    b1, b2, b3 = _bn_struct(bn_exp), _bn_struct(bn_dw), _bn_struct(bn_proj)
    p75, p53 = _transform_ptrs(m75, m53)
    P = n * hw
    _call('mbconv expand 64->%d planar' % mid, 2.0 * P * 64 * mid, P * (64 + mid) * 2,
          lambda: B.check(L.ofa_expand_planar_fwd(x.data_ptr(), t1.data_ptr(), we.data_ptr(), n, hw, mid, xdt, dt,
                                                  byref(b1), act, st)))
    _call('dw%dx%d C%d planar' % (ks, ks, mid), 2.0 * P * mid * ks * ks, 2 * P * mid * 2,
          lambda: B.check(L.ofa_dw_planar_fwd(t1.data_ptr(), t2.data_ptr(), n, mid, h, w, B.fptr(w_dw), w_dw.shape[-1],
                                              p75, p53, int(bool(transform_on)), ks, dt, byref(b2), act, st)))
    _call('mbconv project %d->64 planar' % mid, 2.0 * P * 64 * mid, P * (mid + 64 + (64 if add_residual else 0)) * 2,
          lambda: B.check(L.ofa_project_planar_fwd(t2.data_ptr(), x.data_ptr() if add_residual else None, y.data_ptr(),
                                                   wp.data_ptr(), n, hw, mid, xdt, dt, byref(b3), st)))
    return y


def inference_mode_active(module):
    """The fused single-kernel path is taken when nothing needs autograd and BN uses running stats."""
    return (not module.training) and (not torch.is_grad_enabled())
