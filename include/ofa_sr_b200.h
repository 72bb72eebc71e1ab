/*
 * ofa_sr_b200.h — C ABI of the B200 (sm_100a) elastic-MBConv super-resolution hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI: its "operator API" is a
 * set of Python nn.Modules that call ATen.  Each entry point below therefore cites the reference
 * call site (path:line under the reference tree) whose arithmetic it replaces; the Python host
 * package `ofa_b200` binds them with ctypes (INTEGRATION.md shows the stub a maintainer of the
 * reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless a name says host;
 *   - activations are addressed through explicit element strides (sn, sc, sh, sw), so NCHW,
 *     NHWC (torch channels_last) and channel-sliced views all work without copies; the tensor-core
 *     paths additionally require NHWC-dense 16-bit (bf16 or fp16) and say so;
 *   - dtype codes: OFA_F32 = 0, OFA_BF16 = 1, OFA_F16 = 2 (accumulation is always fp32);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     keeps no global mutable state besides a per-thread error string / launch counter and a
 *     per-device cache of device attributes (nn.DataParallel calls in from one thread per GPU);
 *   - return value: 0 = OFA_OK, otherwise an OFA_ERR_* code; ofa_last_error() gives the text.
 *     The Python layer turns non-zero into RuntimeError (the reference's error style is Python
 *     assert / ValueError: ofa/utils.py:217-218,306).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns OFA_ERR_CUDA.
 */
#ifndef OFA_SR_B200_H_
#define OFA_SR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFA_OK 0
#define OFA_ERR_ARG 1
#define OFA_ERR_CUDA 2
#define OFA_ERR_UNSUPPORTED 3

#define OFA_F32 0
#define OFA_BF16 1
#define OFA_F16 2 /* IEEE half: 16-bit activation storage with 3 more mantissa bits than bf16 */
#define OFA_U8 3  /* OUTPUT tensors of ofa_conv_fwd only (thin-output kernel / CUDA-core kernel): the image as the
                   * reference's consumer makes it, uint8 = round_half_even(clamp(y, 0, 1) * 255) -- tensor2img_np,
                   * sr_run_manager.py:567-597 -- so a 4K frame leaves the device as 25 MB instead of 100 MB */

/* activation codes (ofa/utils.py:242-314 build_activation) */
#define OFA_ACT_NONE 0
#define OFA_ACT_RELU6 1
#define OFA_ACT_HSWISH 2
#define OFA_ACT_RELU 3
#define OFA_ACT_HSIGMOID 4 /* relu6(x + 3) / 6 — the gate of the squeeze-and-excite module (ofa/utils.py:345-352) */

/* store modes of a ConvLayer's third op (ofa/layers.py:94-98 + ofa/utils.py:259-260,383-410) */
#define OFA_STORE_PLAIN 0
#define OFA_STORE_PIXELSHUFFLE2 1   /* out[n, c, 2h+i, 2w+j] = conv[n, 4c+2i+j, h, w] */
#define OFA_STORE_PIXELUNSHUFFLE2 2 /* out[n, 4c+2y+x, h, w] = conv[n, c, 2h+y, 2w+x] */

/* implementation selectors (testing / profiling); 0 picks the fastest valid one */
#define OFA_IMPL_AUTO 0
#define OFA_IMPL_SIMT 1 /* CUDA-core kernels: any layout, fp32 / bf16 / fp16 I/O                */
#define OFA_IMPL_FAST 2 /* TMA halo tiles (depthwise) / tcgen05+TMEM implicit GEMM (dense conv)  */
#define OFA_IMPL_NHWC 3 /* ofa_mbconv_fwd only: the three NHWC kernels instead of the planar path */
#define OFA_IMPL_BAND 4 /* ofa_mbconv_fwd only: force the single-launch band-scheduled block (L2-resident ring)   */
#define OFA_IMPL_PLANAR3 5 /* ofa_mbconv_fwd only: force the three stand-alone planar kernels (HBM intermediates) */

/* A 4-D activation view: element (n, c, h, w) lives at ptr + n*sn + c*sc + h*sh + w*sw (elements). */
typedef struct OfaTensor4 {
  void* ptr;
  int32_t dtype; /* OFA_F32 | OFA_BF16 */
  int32_t n, c, h, w;
  int64_t sn, sc, sh, sw;
} OfaTensor4;

/* Per-output-channel affine + activation + residual applied by every forward kernel's epilogue.
 * Inference BN folding (dynamic_op.py:148-167 with bn.training == False; layers.py:46-50) is done
 * inside the kernel: scale = gamma * rsqrt(var + eps), shift = beta - mean * scale, all read from
 * the FULL supernet-width arrays (the active slice is always the prefix [:C], dynamic_op.py:163-165).
 * Any of gamma/beta/mean/var may be NULL (treated as 1/0/0/1 and eps ignored when var is NULL). */
typedef struct OfaEpilogue {
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* var;
  float eps;
  int32_t act;               /* OFA_ACT_* */
  const OfaTensor4* residual; /* optional tensor added AFTER affine+act (proxyless_nets.py:50
                                 residual, ofa_mbs4.py:159 long skip); indexed like the output y */
} OfaEpilogue;

/* ---------------------------------------------------------------------------------------------
 * Library / device
 * ------------------------------------------------------------------------------------------- */
int ofa_version(void);
const char* ofa_last_error(void);
/* host query; fills SM count and compute capability of the current device */
int ofa_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);
/* number of kernels this library has launched on this thread since the last reset (bench.py's
 * gpu_launches counter) */
int64_t ofa_launch_count(void);
void ofa_launch_count_reset(void);

/* ---------------------------------------------------------------------------------------------
 * (a1) DynamicSeparableConv2d.get_active_filter — dynamic_op.py:46-71
 *   w7   [Cmax, kmax*kmax] fp32 (the nn.Conv2d depthwise weight [Cmax,1,kmax,kmax], contiguous)
 *   m75  [25,25] / m53 [9,9] fp32 or NULL; transform_on mirrors KERNEL_TRANSFORM_MODE is not None
 *   out  [C, ks*ks] fp32, contiguous
 * ------------------------------------------------------------------------------------------- */
int ofa_dw_active_filter(const float* w7, int32_t kmax, const float* m75, const float* m53,
                         int32_t transform_on, int32_t ks, int32_t C, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (a2 [+a5 eval, +a6]) DynamicSeparableConv2d.forward — dynamic_op.py:73-84
 *   depthwise ks x ks, stride 1, dilation 1, pad ks/2, groups = C, filter derived on the fly from
 *   (w7, m75, m53) exactly as (a1).  x and y have C channels, same N/H/W.  epi may be NULL.
 *   OFA_IMPL_FAST requires NHWC-dense bf16 x and y and C % 64 == 0 (the planar tensor-core depthwise
 *   is ofa_dw_planar_fwd below).
 * ------------------------------------------------------------------------------------------- */
int ofa_dw_fwd(const OfaTensor4* x, const OfaTensor4* y, const float* w7, int32_t kmax,
               const float* m75, const float* m53, int32_t transform_on, int32_t ks,
               const OfaEpilogue* epi, int32_t impl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (a3, a4) DynamicPointConv2d.forward — dynamic_op.py:104-112  (1x1, channel-sliced weight)
 * (a9..a11) ConvLayer = conv k x k -> BN -> {none | PixelShuffle(2) | PixelUnshuffle(2)}
 *            — layers.py:94-98,120-151; utils.py:259-260,383-410
 *   w      fp32 weight addressed as w[o*w_so + i*w_si + ky*w_sh + kx*w_sw]: pass the FULL supernet
 *          parameter with its strides and the active Cin/Cout — the slice W[:Cout,:Cin] is never copied
 *   w_bf16 optional packed 16-bit copy [ks*ks][cout_pad][cin_pad] in the SAME format as x (bf16 or fp16),
 *          made by ofa_pack_weight_bf16 / ofa_pack_weight_16
 *          (required by OFA_IMPL_FAST; a derived cache owned by the caller)
 *   y      describes the tensor actually written: for PIXELSHUFFLE2 it has c = Cout/4, h = 2H, w = 2W;
 *          for PIXELUNSHUFFLE2 c = 4*Cout, h = H/2, w = W/2.  epi.residual is indexed like y.
 *          The per-channel affine of epi is indexed by the CONV output channel (BN sits before the
 *          shuffle in the reference, layers.py:94-98).
 * ------------------------------------------------------------------------------------------- */
typedef struct OfaConvArgs {
  OfaTensor4 x;
  OfaTensor4 y;
  const float* w;
  int64_t w_so, w_si, w_sh, w_sw;
  const void* w_bf16;
  int32_t cin_pad, cout_pad; /* padded dims of w_bf16 */
  int32_t cin, cout, ks;
  int32_t flip;  /* 1: use w[.., ks-1-ky, ks-1-kx] (the data-gradient correlation) */
  int32_t store; /* OFA_STORE_* */
  OfaEpilogue epi;
} OfaConvArgs;

int ofa_conv_fwd(const OfaConvArgs* a, int32_t impl, void* stream);
/* named aliases of the same entry point, kept for the reference-facing vocabulary */
int ofa_pw_fwd(const OfaConvArgs* a, int32_t impl, void* stream);       /* requires ks == 1 */
int ofa_conv_kxk_fwd(const OfaConvArgs* a, int32_t impl, void* stream); /* any odd ks        */

/* fp32 strided weight (as above) -> bf16 [ks*ks][cout_pad][cin_pad], zero padded.  `store` is the
 * store mode of the layer the pack is for: PixelShuffle layers are packed sub-pixel major (row
 * s*(cout/4) + c holds conv output channel 4c + s) so one sub-pixel's channels are contiguous. */
int ofa_pack_weight_bf16(const float* w, int64_t w_so, int64_t w_si, int64_t w_sh, int64_t w_sw,
                         int32_t cin, int32_t cout, int32_t ks, int32_t cin_pad, int32_t cout_pad,
                         int32_t store, void* out_bf16, void* stream);
/* same, into `dtype` = OFA_BF16 | OFA_F16: the packed copy handed to OFA_IMPL_FAST must have the 16-bit
 * format of the activation tensor x (tcgen05 kind::f16 multiplies like formats) */
int ofa_pack_weight_16(const float* w, int64_t w_so, int64_t w_si, int64_t w_sh, int64_t w_sw,
                       int32_t cin, int32_t cout, int32_t ks, int32_t cin_pad, int32_t cout_pad,
                       int32_t store, int32_t dtype, void* out, void* stream);

/* Many packs in ONE launch: `jobs_device` is a DEVICE array of `njobs` OfaPackJob records (each is the argument
 * list of ofa_pack_weight_16).  The drop-in modules re-derive every 16-bit weight copy a forward pass uses from
 * the fp32 masters at the start of that pass with one such launch, so a copy can never be stale -- whatever
 * wrote the master (optimizer step, load_state_dict, `w.data.copy_()`, ofa/utils.py:134-155 init_model,
 * elastic_nn/utils.py:76-82 adjust_bn_according_to_idx / dynamic_layers.py:156-199 re-sorting). */
typedef struct OfaPackJob {
  const float* w;
  int64_t w_so, w_si, w_sh, w_sw;
  int32_t cin, cout, ks, cin_pad, cout_pad, store, dtype, reserved;
  void* out;
} OfaPackJob;
int ofa_pack_weights_multi(const OfaPackJob* jobs_device, int32_t njobs, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (a5) DynamicBatchNorm2d.bn_forward in TRAINING mode — dynamic_op.py:148-167
 *   stats : per-channel batch mean and BIASED variance over N*H*W (what F.batch_norm normalises with)
 *   update: running = (1-momentum)*running + momentum*{mean, var * P/(P-1)} on the slice [:C]; when
 *           num_batches_tracked (DEVICE int64, may be NULL) is given it is incremented in the same launch
 *           (`bn.num_batches_tracked += 1`, dynamic_op.py:156)
 *   apply : y = act(gamma*(x-mean)*rsqrt(var+eps)+beta) is ofa_affine_act with mean/var = batch stats
 * ------------------------------------------------------------------------------------------- */
int ofa_bn_stats(const OfaTensor4* x, float* mean, float* var_biased, void* stream);
/* The whole training-mode DynamicBatchNorm2d.bn_forward (+ the activation / residual that follows it) in ONE call and three
 * launches: batch statistics (split reduction), finalize -- which also updates the running-statistics slice and bumps
 * num_batches_tracked when running_mean is given and momentum != 0 --, then y = act(BN(x)) [+ residual].
 * batch_mean / batch_var [C] receive the statistics used (the backward needs them). */
int ofa_bn_train_fwd(const OfaTensor4* x, const OfaTensor4* y, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int32_t act, const OfaTensor4* residual,
                     float* batch_mean, float* batch_var, int64_t* num_batches_tracked, void* stream);
int ofa_bn_update_running(const float* mean, const float* var_biased, int64_t count,
                          float* running_mean, float* running_var, float momentum, int32_t C,
                          int64_t* num_batches_tracked, void* stream);
/* (a5 eval, a6, a8, a10, a11) stand-alone per-channel affine + activation (+ residual) with an
 * optional PixelShuffle / PixelUnshuffle store:  y = store(act(affine(x))) + res */
int ofa_affine_act(const OfaTensor4* x, const OfaTensor4* y, const OfaEpilogue* epi, int32_t store,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * (a7 + a8) DynamicMBConvLayer.forward + MobileInvertedResidualBlock — dynamic_layers.py:70-84,
 * proxyless_nets.py:44-51, inference mode (BN folded): expand 1x1 -> BN -> ReLU6 -> elastic
 * depthwise -> BN -> ReLU6 -> project 1x1 -> BN (+ x).  NHWC-dense x / y with `cin` channels, both bf16
 * or both fp16.
 *   w_exp  [Mmax, cin_max] fp32 (full), w_proj [cout_max, Mmax] fp32 (full); `mid` active channels
 *   bn_*   the three BatchNorm parameter sets (full width; prefixes are used)
 *   ws     caller-owned scratch of ofa_mbconv_workspace_bytes(...) bytes
 * ------------------------------------------------------------------------------------------- */
typedef struct OfaBn {
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* var;
  float eps;
} OfaBn;

typedef struct OfaMBConvArgs {
  OfaTensor4 x;
  OfaTensor4 y;
  const float* w_exp;
  int64_t w_exp_so, w_exp_si;
  const float* w_dw; /* [Mmax, kmax*kmax] */
  int32_t kmax;
  const float* m75;
  const float* m53;
  int32_t transform_on;
  const float* w_proj;
  int64_t w_proj_so, w_proj_si;
  int32_t cin, mid, cout, ks;
  int32_t act; /* activation after expand and depthwise (OFA_ACT_*) */
  OfaBn bn_exp, bn_dw, bn_proj;
  int32_t add_residual; /* 1: y = block(x) + x */
  void* ws;
  int64_t ws_bytes;
  int32_t mid_dtype; /* storage of the two expanded intermediates on the planar tensor-core path:
                        0 = default (OFA_F16: they are ReLU6-clamped, fp16 keeps 3 more mantissa bits),
                        OFA_BF16 or OFA_F16 */
  const void* w_exp_packed;  /* optional (both or neither): the block's weights already packed for the planar path by */
  const void* w_proj_packed; /* ofa_mbconv_pack_weights with the same trunk / mid formats; NULL = pack per call       */
} OfaMBConvArgs;

int64_t ofa_mbconv_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t cin, int32_t mid,
                                   int32_t cout);
/* Frame-sized inputs (>= 128 x 448 pixels) run the planar path as ONE launch (mbconv_band.cu): expand, depthwise and
 * project execute concurrently as role-specialised persistent CTAs, handing 128 x 112-pixel regions to each other
 * through ring buffers small enough to stay in the 126 MB L2, so the two expanded intermediates never reach HBM
 * (dynamic_layers.py:70-84: the reference materialises both, at full size).  OFA_IMPL_BAND forces it on any supported
 * shape, OFA_IMPL_PLANAR3 forces the three stand-alone kernels.
 * impl: OFA_IMPL_AUTO runs the planar tcgen05 path (expand -> Toeplitz depthwise -> project on channel-planar 16-bit
 * intermediates) when cin = cout = 64, mid % 64 == 0, W % 8 == 0 AND the planes have H * W >= 2304 pixels (48 x 48)
 * filling >= 25 % of the 128 (64) x 112-pixel depthwise tiles: the depthwise rebuilds its filter matrices per channel
 * plane, so batches of small patches run faster on the three NHWC kernels, which AUTO picks otherwise.
 * OFA_IMPL_FAST forces the planar path whenever it is supported, OFA_IMPL_NHWC the NHWC kernels, OFA_IMPL_SIMT the
 * exact CUDA-core kernels. */
int ofa_mbconv_fwd(const OfaMBConvArgs* a, int32_t impl, void* stream);

/* The three stages of ofa_mbconv_fwd's planar tcgen05 path (cin = cout = 64, mid % 64 == 0, W % 8 == 0),
 * exported so that tests and profilers can drive them one by one.  "planar" = [N][C][H*W] dense, 16-bit
 * (`dtype` = OFA_BF16 | OFA_F16); "nhwc" = [N][H*W][64] dense 16-bit (`trunk_dtype` = OFA_BF16 | OFA_F16).
 *   pack    : w_exp[:mid,:64] -> `trunk_dtype` [ceil(mid/128)*128][64], w_proj[:64,:mid] -> `dtype` [64][mid]
 *   expand  : (a3 + a5 + a6) 1x1 64 -> mid, folded BN, activation            nhwc   -> planar
 *   dw      : (a1 + a2 + a5 + a6) elastic depthwise ks x ks, folded BN, act   planar -> planar
 *   project : (a4 + a5 + a8) 1x1 mid -> 64, folded BN, + residual (or NULL)   planar -> nhwc            */
int ofa_mbconv_pack_weights(const float* w_exp, int64_t w_exp_so, int64_t w_exp_si, const float* w_proj,
                            int64_t w_proj_so, int64_t w_proj_si, int32_t mid, int32_t trunk_dtype,
                            int32_t dtype, void* wexp_packed, void* wproj_packed, void* stream);
int ofa_expand_planar_fwd(const void* x_nhwc, void* y_planar, const void* wexp_packed, int32_t n, int32_t hw,
                          int32_t mid, int32_t trunk_dtype, int32_t dtype, const OfaBn* bn, int32_t act,
                          void* stream);
int ofa_dw_planar_fwd(const void* x_planar, void* y_planar, int32_t n, int32_t c, int32_t h, int32_t w,
                      const float* w7, int32_t kmax, const float* m75, const float* m53, int32_t transform_on,
                      int32_t ks, int32_t dtype, const OfaBn* bn, int32_t act, void* stream);
int ofa_project_planar_fwd(const void* x_planar, const void* res_nhwc, void* y_nhwc, const void* wproj_packed,
                           int32_t n, int32_t hw, int32_t mid, int32_t trunk_dtype, int32_t dtype,
                           const OfaBn* bn, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (a14) backward of the path — autograd of dynamic_op.py:73-84,104-112,148-167
 * ------------------------------------------------------------------------------------------- */
/* dX of the depthwise conv = correlation of dY with the 180-degree rotated active filter */
int ofa_dw_bwd_data(const OfaTensor4* dy, const OfaTensor4* dx, const float* w7, int32_t kmax,
                    const float* m75, const float* m53, int32_t transform_on, int32_t ks,
                    void* stream);
/* dW_active [C, ks*ks] fp32 (overwritten) */
int ofa_dw_bwd_filter(const OfaTensor4* x, const OfaTensor4* dy, int32_t ks, float* dw_active,
                      void* stream);
/* chain rule through (a1): dW_active -> accumulates (+=) into dw7 [Cmax,kmax*kmax], dm75, dm53 */
int ofa_dw_active_filter_bwd(const float* w7, int32_t kmax, const float* m75, const float* m53,
                             int32_t transform_on, int32_t ks, int32_t C, const float* dw_active,
                             float* dw7, float* dm75, float* dm53, void* stream);
/* dense conv backward: dX (full correlation with W) and dW (accumulated += into the strided slice) */
int ofa_conv_bwd_data(const OfaTensor4* dy, const OfaTensor4* dx, const float* w, int64_t w_so,
                      int64_t w_si, int64_t w_sh, int64_t w_sw, int32_t cin, int32_t cout,
                      int32_t ks, void* stream);
int ofa_conv_bwd_weight(const OfaTensor4* x, const OfaTensor4* dy, float* dw, int64_t w_so,
                        int64_t w_si, int64_t w_sh, int64_t w_sw, int32_t cin, int32_t cout,
                        int32_t ks, void* stream);
/* BN (+ activation) backward, x = the BN input, dy = gradient of act(BN(x)).
 *   reduce: dz = dy * act'(z), z recomputed from x; sum_dz[c], sum_dz_xhat[c] (overwritten)
 *           (dgamma = sum_dz_xhat, dbeta = sum_dz)
 *   apply : training: dx = gamma*rstd*(dz - sum_dz/P - xhat*sum_dz_xhat/P)
 *           eval    : dx = gamma*rstd*dz   (sum pointers may be NULL)                            */
int ofa_bn_bwd_reduce(const OfaTensor4* x, const OfaTensor4* dy, const float* gamma,
                      const float* beta, const float* mean, const float* var, float eps,
                      int32_t act, float* sum_dz, float* sum_dz_xhat, void* stream);
int ofa_bn_bwd_apply(const OfaTensor4* x, const OfaTensor4* dy, const OfaTensor4* dx,
                     const float* gamma, const float* beta, const float* mean, const float* var,
                     float eps, int32_t act, int32_t training, const float* sum_dz,
                     const float* sum_dz_xhat, void* stream);
/* ofa_bn_bwd_reduce followed (dx != NULL) by ofa_bn_bwd_apply in one call */
int ofa_bn_train_bwd(const OfaTensor4* x, const OfaTensor4* dy, const OfaTensor4* dx, const float* gamma,
                     const float* beta, const float* mean, const float* var, float eps, int32_t act, int32_t training,
                     float* sum_dz, float* sum_dz_xhat, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (a7 + a8 + a14) ONE MBConv block of the training step as one call each way — DynamicMBConvLayer.forward in train
 * mode (dynamic_layers.py:70-84: expand 1x1 -> DynamicBatchNorm2d (batch statistics, running-slice update,
 * num_batches_tracked, dynamic_op.py:148-167) -> act -> elastic depthwise -> BN -> act -> project 1x1 -> BN) inside
 * MobileInvertedResidualBlock (proxyless_nets.py:44-51: + x), and its autograd.  Same kernels, order and results as the
 * layer-by-layer entry points above (ofa_conv_fwd / ofa_bn_train_fwd / ofa_dw_fwd ... ofa_conv_bwd_weight); what it
 * removes is host time: the eager progressive-shrinking step (progressive_shrinking.py:158-203) was bounded by the
 * caller's launch loop.  16-bit dense NHWC activations, 64-channel trunk (cin = cout = 64), mid % 64 == 0, mid <= 384.
 *   ws       caller-owned, 256-byte aligned, ofa_mbconv_train_workspace_bytes(...) bytes; the forward leaves the five
 *            intermediates and the batch statistics the backward needs in it (keep it alive and unchanged until then)
 *   BN       gamma / beta / running_* are the FULL-width arrays (prefix [:C] used); running_mean == NULL or
 *            momentum == 0 skips the running update and the counter bump
 *   backward dy, dx: [N, H, W, 64] like x.  Gradient arrays are full parameter size and ZERO-FILLED by the caller:
 *            dw_exp / dw_proj take the parameter's strides, dgamma / dbeta prefixes are overwritten, dw_dw / dm75 / dm53
 *            are accumulated into (dm75 / dm53 may be NULL when the active kernel size does not use the matrix).
 *            With add_residual the identity branch's gradient is included in dx.
 * ------------------------------------------------------------------------------------------- */
typedef struct OfaBnTrain {
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  int64_t* num_batches_tracked;
  float momentum;
  float eps;
} OfaBnTrain;

typedef struct OfaMBConvTrainArgs {
  const void* x; /* [N, H, W, cin]  */
  void* y;       /* [N, H, W, cout] (forward output; unused by the backward) */
  int32_t dtype; /* OFA_BF16 | OFA_F16 */
  int32_t n, h, w;
  int32_t cin, mid, cout, ks, kmax, transform_on;
  int32_t act;          /* activation after BN1 and BN2 (OFA_ACT_*) */
  int32_t add_residual; /* 1: y = block(x) + x */
  const float* w_exp;   /* [Mmax, cin_max, 1, 1] fp32 master, slice [:mid, :cin] read in place */
  int64_t w_exp_so, w_exp_si;
  const float* w_dw; /* [Mmax, kmax * kmax] */
  const float* m75;
  const float* m53;
  const float* w_proj; /* [cout_max, Mmax, 1, 1] */
  int64_t w_proj_so, w_proj_si;
  OfaBnTrain bn_exp, bn_dw, bn_proj;
  void* ws;
  int64_t ws_bytes;
} OfaMBConvTrainArgs;

typedef struct OfaMBConvTrainGrads {
  float* dw_exp;
  float* dw_dw;
  float* dm75;
  float* dm53;
  float* dw_proj;
  float* dgamma[3]; /* expand, depthwise, project BatchNorm */
  float* dbeta[3];
} OfaMBConvTrainGrads;

/* The three weight-gradient computations of a block do not feed its data-gradient chain: with mode 1 ofa_mbconv_train_bwd
 * runs them on a second, lower-priority stream beside it (fork / join inside the call, so the caller's stream semantics
 * are unchanged).  mode 0 (library default): everything on the caller's stream.  Returns the previous mode. */
int ofa_train_side_mode(int32_t mode);
int64_t ofa_mbconv_train_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t cin, int32_t mid, int32_t cout);
int ofa_mbconv_train_fwd(const OfaMBConvTrainArgs* a, void* stream);
int ofa_mbconv_train_bwd(const OfaMBConvTrainArgs* a, const void* dy, void* dx, const OfaMBConvTrainGrads* g,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * (evaluation side) the validate metric on the device —
 * psnr(rgb2y(tensor2img_np(a)), rgb2y(tensor2img_np(b))): sr_run_manager.py:364,496,567-597, ofa/utils.py:27-34.
 *   a, b           [N,3,H,W] images (any strides, fp32 or 16-bit), values clamped to [0,1] as the reference does
 *   sse_per_image  DEVICE int64[N] (overwritten): sum over pixels of (Y_a - Y_b)^2 with Y the BT.601 luma of the
 *                  uint8-rounded image — an exact integer; the host turns it into PSNR (ofa_b200.metrics.psnr_y
 *                  also reproduces the reference's make_grid padding quirk for N > 1)
 * ------------------------------------------------------------------------------------------- */
int ofa_psnr_y_sse(const OfaTensor4* a, const OfaTensor4* b, int64_t* sse_per_image, void* stream);

/* ---------------------------------------------------------------------------------------------
 *  multi-tensor Adam — torch.optim.Adam(net_params, init_lr) with the per-group L2 weight decay
 * of sr_run_manager.py:115-133,180-185 ('bn#bias' keys: 0), betas / eps as given, no amsgrad.
 *   table_dev   DEVICE array of n_tensors descriptors (parameter, both moments, size, the tensor's weight decay)
 *   chunks_dev  DEVICE int32 pairs (tensor index, chunk index): one CUDA block per 2048-element chunk
 *   grads_dev   DEVICE array of n_tensors gradient pointers for THIS step; NULL = the tensor had no gradient (its
 *               block is outside the sampled sub-network): it is skipped, step counter and moments untouched,
 *               exactly as optimizer.step() skips p.grad is None
 *   steps_dev   DEVICE int32[n_tensors] per-tensor step counters (start at 0), bumped for the active tensors
 * ------------------------------------------------------------------------------------------- */
typedef struct OfaAdamTensor {
  float* p;
  float* m;
  float* v;
  int64_t numel;
  float weight_decay;
  int32_t reserved;
} OfaAdamTensor;

int ofa_adam_step(const OfaAdamTensor* table_dev, const int32_t* chunks_dev, int32_t n_tensors, int32_t n_chunks,
                  const float* const* grads_dev, int32_t* steps_dev, float lr, float beta1, float beta2,
                  float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (training side) SR data preparation on the device — what the reference's data-loader workers
 * do per sample with Pillow / torchvision: ofa/imagenet_codebase/data_providers/div2k_setxx.py:166-171 (RandomCrop,
 * RandomHorizontalFlip, RandomRotation), :288-298 (L2 / L4 = Scale(1/2), Scale(1/4) with Image.BICUBIC, ToTensor),
 * :355-380 (Scale).  Bit-exact against Pillow's libImaging (Resample.c, Geometry.c) and torchvision's ToTensor.
 * uint8 images are pixel-interleaved RGB, [N][H][W][3]; fp32 outputs are planar [N][3][H][W] = value / 255.
 *
 * ofa_resample_ksize / ofa_resample_build_table: HOST functions; the per-output-index window (xmin, count) and the
 *   22-bit fixed-point bicubic coefficients Pillow's precompute_coeffs + normalize_coeffs_8bpc produce for resizing
 *   an axis of in_size to out_size.  bounds_host: int32[out_size][2], kk_host: int32[out_size][ksize].  The caller
 *   copies them to the device once per (in_size, out_size) and owns them (the library keeps no table cache).
 * ofa_bicubic_resize_u8: img.resize((out_w, out_h), Image.BICUBIC) for a batch: horizontal pass into `tmp`
 *   ([N][H][out_w][3] uint8 scratch), vertical pass into out_u8 and / or out_f32 (either may be NULL).
 * ofa_sr_augment_u8: one size x size patch per sample: crop at (row i, column j) of the sample's source image
 *   (src + n * sample_stride bytes, [h][w][3]), optional left-right flip, rotation.
 *   params: DEVICE int32[N][10] = { i, j, flip, mode, a0, a1, a2, a3, a4, a5 }; mode 0 none, 1 = 180 deg, 2 / 3 =
 *   90 / 270 deg (Pillow's transposes, square patches), 4 = affine nearest-neighbour walk with the 16.16
 *   fixed-point inverse matrix of Geometry.c (the host computes it: ofa_b200/data.py rotation_params).
 * ------------------------------------------------------------------------------------------- */
int32_t ofa_resample_ksize(int32_t in_size, int32_t out_size);
int ofa_resample_build_table(int32_t in_size, int32_t out_size, int32_t* bounds_host, int32_t* kk_host);
int ofa_bicubic_resize_u8(const uint8_t* src, int32_t n, int32_t h, int32_t w, int32_t out_h, int32_t out_w,
                          const int32_t* bounds_h, const int32_t* kk_h, int32_t ksize_h, const int32_t* bounds_v,
                          const int32_t* kk_v, int32_t ksize_v, uint8_t* tmp, uint8_t* out_u8, float* out_f32,
                          void* stream);
int ofa_sr_augment_u8(const uint8_t* src, int64_t sample_stride, int32_t n, int32_t h, int32_t w,
                      const int32_t* params, int32_t size, uint8_t* out_u8, float* out_f32, void* stream);

/* ---------------------------------------------------------------------------------------------
 *  the MobileNetV3 flavour of the elastic modules — ops the SR nets do not exercise.
 * Exact fp32 CUDA-core kernels; row-major fp32 matrices with explicit leading dimensions (elements).
 *
 * ofa_linear_fwd        DynamicLinear.forward (dynamic_op.py:115-136) and the two 1x1 convs of DynamicSE on the pooled
 *                       [N, C, 1, 1] tensor (:175-200): y[n, o] = act(bias[o] + sum_i x[n, i] * w[o * ldw + i]); the
 *                       active slice w[:out, :in] is addressed in place through ldw (the reference copies it).
 * ofa_act_bwd_from_output   dz = dy * act'(z), derivative taken from the OUTPUT (ReLU / ReLU6 / h-sigmoid)
 * ofa_linear_bwd_data   dx[n, i] = sum_o dz[n, o] * w[o * ldw + i]
 * ofa_linear_bwd_weight dw[o * lddw + i] = sum_n dz[n, o] * x[n, i] (overwrites the active slice), db[o] = sum_n dz
 * ofa_plane_mean        pooled[n * C + c] = mean over (h, w) of x — SEModule's x.mean(3).mean(2) (ofa/utils.py:371)
 * ofa_plane_dot         out[n * C + c] = sum over (h, w) of x * dy — gradient of the gate in y = x * s
 * ofa_channel_scale     y[n, c, h, w] = x[n, c, h, w] * s[n * C + c] (+ add[n * C + c] when not NULL) — `x * y` of
 *                       SEModule.forward (:373), and with (dy, s, dpooled / HW) its input gradient
 * ofa_dw_strided_*      DynamicSeparableConv2d.forward with stride > 1 (dynamic_op.py:73-84; same padding ks // 2,
 *                       output (H - 1) / stride + 1) on the explicit active filter [C][ks * ks] that
 *                       ofa_dw_active_filter produced; the filter gradient is overwritten (not accumulated) and
 *                       continues through ofa_dw_active_filter_bwd.
 * ------------------------------------------------------------------------------------------- */
int ofa_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int32_t n,
                   int32_t in_features, int32_t out_features, int32_t act, float* y, int64_t ldy, void* stream);
int ofa_act_bwd_from_output(const float* dy, const float* y, int32_t act, float* dz, int64_t total, void* stream);
int ofa_linear_bwd_data(const float* dz, int64_t lddz, const float* w, int64_t ldw, int32_t n, int32_t in_features,
                        int32_t out_features, float* dx, int64_t lddx, void* stream);
int ofa_linear_bwd_weight(const float* dz, int64_t lddz, const float* x, int64_t ldx, int32_t n, int32_t in_features,
                          int32_t out_features, float* dw, int64_t lddw, float* db, void* stream);
int ofa_plane_mean(const OfaTensor4* x, float* pooled, void* stream);
int ofa_plane_dot(const OfaTensor4* x, const OfaTensor4* dy, float* out, void* stream);
int ofa_channel_scale(const OfaTensor4* x, const OfaTensor4* y, const float* s, const float* add, void* stream);
int ofa_dw_strided_fwd(const OfaTensor4* x, const OfaTensor4* y, const float* filt, int32_t ks, int32_t stride,
                       const OfaEpilogue* epi, void* stream);
int ofa_dw_strided_bwd_data(const OfaTensor4* dy, const OfaTensor4* dx, const float* filt, int32_t ks, int32_t stride,
                            void* stream);
int ofa_dw_strided_bwd_filter(const OfaTensor4* x, const OfaTensor4* dy, int32_t ks, int32_t stride, float* dfilt,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OFA_SR_B200_H_ */
