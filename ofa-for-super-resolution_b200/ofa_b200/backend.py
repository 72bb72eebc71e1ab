"""ctypes binding of libofa_sr_b200.so — the C-ABI library declared in include/ofa_sr_b200.h.

PyTorch is used here only for device memory (`tensor.data_ptr()`), strides and the current CUDA
stream; every computation of the hot path happens inside the library's hand-written sm_100a
kernels.  There is NO fallback: a missing library raises at import of the first op, and a call on a
non-CUDA tensor raises RuntimeError.
"""
import ctypes
import os
from ctypes import c_int32, c_int64, c_float, c_void_p, c_char_p, POINTER, Structure, byref

import torch

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib')
LIB_PATH = os.path.join(_LIB_DIR, 'libofa_sr_b200.so')

OFA_F32, OFA_BF16, OFA_F16, OFA_U8 = 0, 1, 2, 3
ACT_NONE, ACT_RELU6, ACT_HSWISH, ACT_RELU, ACT_HSIGMOID = 0, 1, 2, 3, 4
STORE_PLAIN, STORE_PIXELSHUFFLE2, STORE_PIXELUNSHUFFLE2 = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_FAST, IMPL_NHWC, IMPL_BAND, IMPL_PLANAR3 = 0, 1, 2, 3, 4, 5

ACT_CODES = {None: ACT_NONE, 'relu6': ACT_RELU6, 'h_swish': ACT_HSWISH, 'relu': ACT_RELU}


class OfaTensor4(Structure):
    _fields_ = [('ptr', c_void_p), ('dtype', c_int32), ('n', c_int32), ('c', c_int32), ('h', c_int32),
                ('w', c_int32), ('sn', c_int64), ('sc', c_int64), ('sh', c_int64), ('sw', c_int64)]


class OfaEpilogue(Structure):
    _fields_ = [('gamma', c_void_p), ('beta', c_void_p), ('mean', c_void_p), ('var', c_void_p),
                ('eps', c_float), ('act', c_int32), ('residual', POINTER(OfaTensor4))]


class OfaConvArgs(Structure):
    _fields_ = [('x', OfaTensor4), ('y', OfaTensor4), ('w', c_void_p),
                ('w_so', c_int64), ('w_si', c_int64), ('w_sh', c_int64), ('w_sw', c_int64),
                ('w_bf16', c_void_p), ('cin_pad', c_int32), ('cout_pad', c_int32),
                ('cin', c_int32), ('cout', c_int32), ('ks', c_int32), ('flip', c_int32),
                ('store', c_int32), ('epi', OfaEpilogue)]


class OfaBn(Structure):
    _fields_ = [('gamma', c_void_p), ('beta', c_void_p), ('mean', c_void_p), ('var', c_void_p),
                ('eps', c_float)]


class OfaMBConvArgs(Structure):
    _fields_ = [('x', OfaTensor4), ('y', OfaTensor4),
                ('w_exp', c_void_p), ('w_exp_so', c_int64), ('w_exp_si', c_int64),
                ('w_dw', c_void_p), ('kmax', c_int32), ('m75', c_void_p), ('m53', c_void_p),
                ('transform_on', c_int32),
                ('w_proj', c_void_p), ('w_proj_so', c_int64), ('w_proj_si', c_int64),
                ('cin', c_int32), ('mid', c_int32), ('cout', c_int32), ('ks', c_int32), ('act', c_int32),
                ('bn_exp', OfaBn), ('bn_dw', OfaBn), ('bn_proj', OfaBn),
                ('add_residual', c_int32), ('ws', c_void_p), ('ws_bytes', c_int64), ('mid_dtype', c_int32),
                ('w_exp_packed', c_void_p), ('w_proj_packed', c_void_p)]


class OfaBnTrain(Structure):
    _fields_ = [('gamma', c_void_p), ('beta', c_void_p), ('running_mean', c_void_p), ('running_var', c_void_p),
                ('num_batches_tracked', c_void_p), ('momentum', c_float), ('eps', c_float)]


class OfaMBConvTrainArgs(Structure):
    _fields_ = [('x', c_void_p), ('y', c_void_p), ('dtype', c_int32), ('n', c_int32), ('h', c_int32), ('w', c_int32),
                ('cin', c_int32), ('mid', c_int32), ('cout', c_int32), ('ks', c_int32), ('kmax', c_int32),
                ('transform_on', c_int32), ('act', c_int32), ('add_residual', c_int32),
                ('w_exp', c_void_p), ('w_exp_so', c_int64), ('w_exp_si', c_int64),
                ('w_dw', c_void_p), ('m75', c_void_p), ('m53', c_void_p),
                ('w_proj', c_void_p), ('w_proj_so', c_int64), ('w_proj_si', c_int64),
                ('bn_exp', OfaBnTrain), ('bn_dw', OfaBnTrain), ('bn_proj', OfaBnTrain),
                ('ws', c_void_p), ('ws_bytes', c_int64)]


class OfaMBConvTrainGrads(Structure):
    _fields_ = [('dw_exp', c_void_p), ('dw_dw', c_void_p), ('dm75', c_void_p), ('dm53', c_void_p), ('dw_proj', c_void_p),
                ('dgamma', c_void_p * 3), ('dbeta', c_void_p * 3)]


class OfaPackJob(Structure):
    _fields_ = [('w', c_void_p), ('w_so', c_int64), ('w_si', c_int64), ('w_sh', c_int64), ('w_sw', c_int64),
                ('cin', c_int32), ('cout', c_int32), ('ks', c_int32), ('cin_pad', c_int32), ('cout_pad', c_int32),
                ('store', c_int32), ('dtype', c_int32), ('reserved', c_int32), ('out', c_void_p)]


# every symbol include/ofa_sr_b200.h declares: name -> (restype, argtypes)
_T4 = POINTER(OfaTensor4)
_EP = POINTER(OfaEpilogue)
SYMBOLS = {
    'ofa_version': (c_int32, []),
    'ofa_last_error': (c_char_p, []),
    'ofa_device_info': (c_int32, [POINTER(c_int32)] * 3),
    'ofa_launch_count': (c_int64, []),
    'ofa_launch_count_reset': (None, []),
    'ofa_dw_active_filter': (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                       c_void_p, c_void_p]),
    'ofa_dw_fwd': (c_int32, [_T4, _T4, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, _EP,
                             c_int32, c_void_p]),
    'ofa_conv_fwd': (c_int32, [POINTER(OfaConvArgs), c_int32, c_void_p]),
    'ofa_pw_fwd': (c_int32, [POINTER(OfaConvArgs), c_int32, c_void_p]),
    'ofa_conv_kxk_fwd': (c_int32, [POINTER(OfaConvArgs), c_int32, c_void_p]),
    'ofa_pack_weight_bf16': (c_int32, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32,
                                       c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    'ofa_pack_weight_16': (c_int32, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32,
                                     c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    'ofa_pack_weights_multi': (c_int32, [c_void_p, c_int32, c_void_p]),
    'ofa_bn_stats': (c_int32, [_T4, c_void_p, c_void_p, c_void_p]),
    'ofa_bn_train_fwd': (c_int32, [_T4, _T4, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int32, _T4,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    'ofa_bn_train_bwd': (c_int32, [_T4, _T4, _T4, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int32, c_int32,
                                   c_void_p, c_void_p, c_void_p]),
    'ofa_bn_update_running': (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_float, c_int32,
                                        c_void_p, c_void_p]),
    'ofa_affine_act': (c_int32, [_T4, _T4, _EP, c_int32, c_void_p]),
    'ofa_mbconv_workspace_bytes': (c_int64, [c_int32] * 6),
    'ofa_mbconv_fwd': (c_int32, [POINTER(OfaMBConvArgs), c_int32, c_void_p]),
    'ofa_mbconv_pack_weights': (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int32, c_int32,
                                          c_int32, c_void_p, c_void_p, c_void_p]),
    'ofa_expand_planar_fwd': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        POINTER(OfaBn), c_int32, c_void_p]),
    'ofa_dw_planar_fwd': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32,
                                    c_void_p, c_void_p, c_int32, c_int32, c_int32, POINTER(OfaBn), c_int32, c_void_p]),
    'ofa_project_planar_fwd': (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                         c_int32, POINTER(OfaBn), c_void_p]),
    'ofa_train_side_mode': (c_int32, [c_int32]),
    'ofa_mbconv_train_workspace_bytes': (c_int64, [c_int32] * 6),
    'ofa_mbconv_train_fwd': (c_int32, [POINTER(OfaMBConvTrainArgs), c_void_p]),
    'ofa_mbconv_train_bwd': (c_int32, [POINTER(OfaMBConvTrainArgs), c_void_p, c_void_p, POINTER(OfaMBConvTrainGrads),
                                       c_void_p]),
    'ofa_adam_step': (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_float, c_float, c_float,
                                c_float, c_void_p]),
    'ofa_psnr_y_sse': (c_int32, [_T4, _T4, c_void_p, c_void_p]),
    'ofa_dw_bwd_data': (c_int32, [_T4, _T4, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32,
                                  c_void_p]),
    'ofa_dw_bwd_filter': (c_int32, [_T4, _T4, c_int32, c_void_p, c_void_p]),
    'ofa_dw_active_filter_bwd': (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ofa_conv_bwd_data': (c_int32, [_T4, _T4, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32,
                                    c_int32, c_int32, c_void_p]),
    'ofa_conv_bwd_weight': (c_int32, [_T4, _T4, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32,
                                      c_int32, c_int32, c_void_p]),
    'ofa_bn_bwd_reduce': (c_int32, [_T4, _T4, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int32,
                                    c_void_p, c_void_p, c_void_p]),
    'ofa_bn_bwd_apply': (c_int32, [_T4, _T4, _T4, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int32,
                                   c_int32, c_void_p, c_void_p, c_void_p]),
    'ofa_linear_fwd': (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                 c_void_p, c_int64, c_void_p]),
    'ofa_act_bwd_from_output': (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p]),
    'ofa_linear_bwd_data': (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p,
                                      c_int64, c_void_p]),
    'ofa_linear_bwd_weight': (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p,
                                        c_int64, c_void_p, c_void_p]),
    'ofa_plane_mean': (c_int32, [_T4, c_void_p, c_void_p]),
    'ofa_plane_dot': (c_int32, [_T4, _T4, c_void_p, c_void_p]),
    'ofa_channel_scale': (c_int32, [_T4, _T4, c_void_p, c_void_p, c_void_p]),
    'ofa_dw_strided_fwd': (c_int32, [_T4, _T4, c_void_p, c_int32, c_int32, _EP, c_void_p]),
    'ofa_dw_strided_bwd_data': (c_int32, [_T4, _T4, c_void_p, c_int32, c_int32, c_void_p]),
    'ofa_dw_strided_bwd_filter': (c_int32, [_T4, _T4, c_int32, c_int32, c_void_p, c_void_p]),
    'ofa_resample_ksize': (c_int32, [c_int32, c_int32]),
    'ofa_resample_build_table': (c_int32, [c_int32, c_int32, c_void_p, c_void_p]),
    'ofa_bicubic_resize_u8': (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                        c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ofa_sr_augment_u8': (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p,
                                    c_void_p, c_void_p]),
}

_lib = None


def lib():
    """Load the library once; fail loudly when it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libofa_sr_b200.so not found at %s — build it with `python -c "import __graft_entry__ as g; '
                'g.build()"` or `make -C ofa-for-super-resolution_b200/csrc`; there is no fallback path' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().ofa_last_error()
        raise RuntimeError('libofa_sr_b200: %s (code %d)' % (msg.decode() if msg else '?', rc))


def _dtype_code(t):
    if t.dtype == torch.float32:
        return OFA_F32
    if t.dtype == torch.bfloat16:
        return OFA_BF16
    if t.dtype == torch.float16:
        return OFA_F16
    raise RuntimeError('libofa_sr_b200 supports float32, bfloat16 and float16 activations, got %s' % t.dtype)


_DTYPE_CODES = {torch.float32: OFA_F32, torch.bfloat16: OFA_BF16, torch.float16: OFA_F16, torch.uint8: OFA_U8}


def dtype_code(dtype):
    return {torch.float32: OFA_F32, torch.bfloat16: OFA_BF16, torch.float16: OFA_F16}[dtype]


def t4(t):
    """torch [N,C,H,W] tensor (any strides) -> OfaTensor4 view.  The tensor must be on a CUDA device."""
    if not t.is_cuda:
        raise RuntimeError('libofa_sr_b200 has no CPU path: tensor is on %s' % t.device)
    code = _DTYPE_CODES.get(t.dtype)
    if code is None:
        raise RuntimeError('libofa_sr_b200 supports float32, bfloat16 and float16 activations, got %s' % t.dtype)
    n, c, h, w = t.shape                     # raises for anything that is not 4-D
    sn, sc, sh, sw = t.stride()
    return OfaTensor4(t.data_ptr(), code, n, c, h, w, sn, sc, sh, sw)


def fptr(t):
    """fp32 device pointer of a parameter / buffer (or NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('libofa_sr_b200 has no CPU path: parameter is on %s' % t.device)
    assert t.dtype == torch.float32 and t.is_contiguous(), 'parameters must be contiguous fp32'
    return t.data_ptr()


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


_raw_device = getattr(torch._C, '_cuda_getDevice', None)


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (the stream every library call launches on).  The raw
    lookup costs ~0.3 us; torch.cuda.current_stream(device).cuda_stream builds a Stream object (~7 us), which at
    ~250 library calls per training step was 13 % of the eager step.

    The library launches on the CURRENT device (kernels, cudaMallocAsync scratch, SM count), so when the tensor lives
    on another device -- `net.to('cuda:1')` while cuda:0 is current; stock torch modules guard against this themselves
    -- the current device is switched to the tensor's before the call (every library call fetches its stream here
    first)."""
    if device is None:
        idx = torch.cuda.current_device()
    else:
        idx = device.index if isinstance(device, torch.device) else int(device)
        if idx is None:
            idx = torch.cuda.current_device()
    cur = _raw_device() if _raw_device is not None else torch.cuda.current_device()
    if idx != cur:
        torch.cuda.set_device(idx)
    if _raw_stream is not None:
        return _raw_stream(idx)
    return torch.cuda.current_stream(device).cuda_stream


def new_nhwc(n, c, h, w, dtype, device):
    """Logical NCHW tensor stored channels-last (the layout every kernel is coalesced for)."""
    return torch.empty((n, c, h, w), dtype=dtype, device=device, memory_format=torch.channels_last)


def epilogue(gamma=None, beta=None, mean=None, var=None, eps=0.0, act=ACT_NONE, residual=None, n_active=None):
    """Build an OfaEpilogue.  Returns (struct, keepalive) — keepalive holds the residual view."""
    res_t4 = t4(residual) if residual is not None else None
    e = OfaEpilogue(fptr(gamma), fptr(beta), fptr(mean), fptr(var), float(eps), int(act),
                    ctypes.pointer(res_t4) if res_t4 is not None else None)
    return e, res_t4


def launch_count():
    return lib().ofa_launch_count()


def launch_count_reset():
    lib().ofa_launch_count_reset()
