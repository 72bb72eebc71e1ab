// Training step of ONE MBConv block (DynamicMBConvLayer.forward in train mode, dynamic_layers.py:70-84, inside
// MobileInvertedResidualBlock, proxyless_nets.py:44-51) as one library call each way.
//
//   forward :  x --1x1 expand--> z1 --BN1(batch stats)+act--> a1 --elastic depthwise--> z2 --BN2+act--> a2
//                --1x1 project--> z3 --BN3 (+ x)--> y
//   backward:  dy -> (dx, dW_exp, dW_dw [+ dM75, dM53], dW_proj, dgamma / dbeta of the three BatchNorms)
//
// The kernels are the ones the layer-by-layer autograd path launches (conv_tc / wgrad_tc on tcgen05, dw_fast,
// dw_bwd_filter_rows, the vec8 BatchNorm kernels), in the same order with the same arguments: outputs, BatchNorm
// gradients and buffers equal that path bit for bit, the atomically summed weight gradients to rounding
// (test_block_training_call_is_bit_identical_to_layerwise_path).  What changes is the HOST side: the eager training
// step was bounded by the Python launch loop (~10 ms of interpreter time per step for ~7.5 ms of kernels,
// tools/hosttime_train.py); a block now costs one autograd node and one ctypes call each way instead of three nodes and
// ~12 calls.  Device-side differences: the block's four weight copies are packed by ONE launch in the forward (the two
// transposed ones are kept in the workspace for the backward); the identity branch's gradient is folded into the fp32
// epilogue of the expand data-gradient conv (dx = conv(dz1, W_exp^T) + dy: one rounding instead of two); the three
// weight-gradient computations can run on a side stream (ofa_train_side_mode).  The five intermediates the backward
// needs live in one caller-owned workspace; backward scratch is stream-ordered (cudaMallocAsync).
#include "ofa_common.cuh"
#include "kernels.h"

#include <stdlib.h>
#include <string.h>

namespace ofa {
namespace {

inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

struct TrainLayout {
  int64_t P;
  int64_t z1, a1, z2, a2, z3;     // byte offsets of the saved activations
  int64_t stats;                  // 2 * (mid + mid + cout) floats: mean1 var1 mean2 var2 mean3 var3
  int64_t wexp, wproj;            // packed 16-bit weights of the two forward convs
  int64_t wexp_t, wproj_t;        // ... and of the two data-gradient convs (transposed slices), packed by the forward
  int64_t total;
};

TrainLayout train_layout(int n, int h, int w, int mid, int cout) {
  TrainLayout L;
  L.P = (int64_t)n * h * w;
  const int64_t midb = align256(L.P * mid * 2), outb = align256(L.P * cout * 2);
  int64_t o = 0;
  L.z1 = o; o += midb;
  L.a1 = o; o += midb;
  L.z2 = o; o += midb;
  L.a2 = o; o += midb;
  L.z3 = o; o += outb;
  L.stats = o; o += align256((int64_t)sizeof(float) * 2 * (2 * mid + cout));
  L.wexp = o; o += align256((int64_t)384 * 64 * 2);
  L.wproj = o; o += align256((int64_t)64 * 384 * 2);
  L.wexp_t = o; o += align256((int64_t)64 * 384 * 2);
  L.wproj_t = o; o += align256((int64_t)384 * 64 * 2);
  L.total = o;
  return L;
}

OfaTensor4 nhwc16(const void* p, int dtype, int n, int c, int h, int w) {
  OfaTensor4 t;
  t.ptr = const_cast<void*>(p);
  t.dtype = dtype;
  t.n = n; t.c = c; t.h = h; t.w = w;
  t.sc = 1; t.sw = c; t.sh = (int64_t)w * c; t.sn = (int64_t)h * w * c;
  return t;
}

int check_args(const OfaMBConvTrainArgs* a, const char* who) {
  OFA_REQUIRE(a != nullptr, "%s: null args", who);
  OFA_REQUIRE(a->dtype == OFA_BF16 || a->dtype == OFA_F16, "%s: activations must be OFA_BF16 or OFA_F16", who);
  OFA_REQUIRE(a->n > 0 && a->h > 0 && a->w > 0, "%s: empty batch", who);
  OFA_REQUIRE(a->cin == 64 && a->cout == 64, "%s: the block path needs a 64-channel trunk (cin = cout = 64)", who);
  OFA_REQUIRE(a->mid % 64 == 0 && a->mid >= 64 && a->mid <= 384, "%s: mid must be a multiple of 64 in [64, 384]", who);
  OFA_REQUIRE(a->ks == 3 || a->ks == 5 || a->ks == 7, "%s: kernel size must be 3, 5 or 7", who);
  OFA_REQUIRE(a->kmax == 3 || a->kmax == 5 || a->kmax == 7, "%s: kmax must be 3, 5 or 7", who);
  OFA_REQUIRE(a->ks <= a->kmax, "%s: active kernel size %d exceeds the stored %d", who, a->ks, a->kmax);
  if (a->transform_on && a->ks < a->kmax) {
    if (a->ks == 5) OFA_REQUIRE(a->m75 != nullptr, "%s: 7to5 matrix required for ks = 5", who);
    if (a->ks == 3) OFA_REQUIRE(a->m53 != nullptr, "%s: a ->3 transform matrix is required for ks = 3", who);
  }
  OFA_REQUIRE(a->act >= OFA_ACT_NONE && a->act <= OFA_ACT_RELU, "%s: bad activation code %d", who, a->act);
  OFA_REQUIRE(a->x && a->y && a->w_exp && a->w_dw && a->w_proj && a->ws, "%s: null pointer", who);
  OFA_REQUIRE(((uintptr_t)a->x & 15) == 0 && ((uintptr_t)a->y & 15) == 0 && ((uintptr_t)a->ws & 255) == 0,
              "%s: x / y must be 16-byte aligned, ws 256-byte aligned", who);
  const OfaBnTrain* bns[3] = {&a->bn_exp, &a->bn_dw, &a->bn_proj};
  for (int i = 0; i < 3; ++i) {
    OFA_REQUIRE(bns[i]->gamma && bns[i]->beta, "%s: BatchNorm %d needs gamma and beta", who, i + 1);
    OFA_REQUIRE((bns[i]->running_mean == nullptr) == (bns[i]->running_var == nullptr), "%s: running_mean / running_var", who);
  }
  const TrainLayout L = train_layout(a->n, a->h, a->w, a->mid, a->cout);
  OFA_REQUIRE(a->ws_bytes >= L.total, "%s: workspace too small (%lld < %lld)", who, (long long)a->ws_bytes,
              (long long)L.total);
  return OFA_OK;
}

// y = conv1x1(x, packed) on the tcgen05 implicit-GEMM kernel, optional residual in the epilogue
int pointwise_tc(const OfaTensor4& x, const OfaTensor4& y, const void* packed, int cin, int cout, const OfaTensor4* res,
                 cudaStream_t st) {
  OfaConvArgs c;
  memset(&c, 0, sizeof(c));
  c.x = x; c.y = y;
  c.w_bf16 = packed;
  c.cin = cin; c.cout = cout; c.ks = 1;
  c.cin_pad = cin; c.cout_pad = (cout + 15) / 16 * 16;
  c.store = OFA_STORE_PLAIN;
  c.epi.act = OFA_ACT_NONE;
  c.epi.residual = res;
  if (!conv_tc_supported(&c)) return fail(OFA_ERR_UNSUPPORTED, "mbconv training block: 1x1 conv %d -> %d not supported by the tensor-core kernel", cin, cout);
  return launch_conv_tc(&c, st);
}

int bn_fwd(const OfaTensor4& z, const OfaTensor4& y, const OfaBnTrain& bn, float* mean, float* var, int act,
           const OfaTensor4* res, cudaStream_t st) {
  const bool upd = bn.running_mean != nullptr && bn.momentum != 0.f;
  int rc = launch_bn_stats_update(make_tv(&z), mean, var, upd ? bn.running_mean : nullptr, upd ? bn.running_var : nullptr,
                                  bn.momentum, upd ? reinterpret_cast<long long*>(bn.num_batches_tracked) : nullptr, st);
  if (rc) return rc;
  OfaEpilogue e;
  memset(&e, 0, sizeof(e));
  e.gamma = bn.gamma; e.beta = bn.beta; e.mean = mean; e.var = var; e.eps = bn.eps; e.act = act; e.residual = res;
  return launch_affine_act(make_tv(&z), make_tv(&y), make_epi(&e), OFA_STORE_PLAIN, st);
}

int bn_bwd(const OfaTensor4& z, const OfaTensor4& dy, const OfaTensor4& dz, const OfaBnTrain& bn, const float* mean,
           const float* var, int act, float* dbeta, float* dgamma, cudaStream_t st) {
  int rc = launch_bn_bwd_reduce(make_tv(&z), make_tv(&dy), bn.gamma, bn.beta, mean, var, bn.eps, act, dbeta, dgamma, st);
  if (rc) return rc;
  return launch_bn_bwd_apply(make_tv(&z), make_tv(&dy), make_tv(&dz), bn.gamma, bn.beta, mean, var, bn.eps, act, 1, dbeta,
                             dgamma, st);
}


// Side stream of the backward: the three weight-gradient computations of a block (project wgrad, depthwise filter
// gradient + transform chain, expand wgrad) do not feed the block's data-gradient chain, so they run on a second,
// lower-priority stream beside it: fork after the BatchNorm backward that produces their dZ, join before the call
// returns (ofa_train_side_mode(1)).  At batch 64 x 24 x 24 every kernel is one latency-bound wave, so the two chains fill
// each other's gaps: max-sub-network step 8.48 -> 8.18 ms.  A deferred join (the main stream never waiting inside the
// backward pass, scratch released on the side stream) was measured and dropped: freeing on another stream than the one
// that allocates defeats the stream-ordered pool's reuse (12-100 ms steps).
struct SideCtx {
  cudaStream_t s;
  cudaEvent_t fork[3], join;
  bool ok;
};

SideCtx g_side[64];
unsigned char g_side_init[64] = {0};
int g_side_mode = 0;

SideCtx* side_ctx() {
  if (!g_side_mode) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_side_init[dev]) {
    g_side_init[dev] = 1;
    SideCtx& c = g_side[dev];
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    c.ok = cudaStreamCreateWithPriority(&c.s, cudaStreamNonBlocking, lo) == cudaSuccess;
    for (int i = 0; i < 3 && c.ok; ++i) c.ok = cudaEventCreateWithFlags(&c.fork[i], cudaEventDisableTiming) == cudaSuccess;
    if (c.ok) c.ok = cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) == cudaSuccess;
    if (!c.ok) cudaGetLastError();
  }
  return g_side[dev].ok ? &g_side[dev] : nullptr;
}

}  // namespace
}  // namespace ofa

using namespace ofa;

extern "C" {

int ofa_train_side_mode(int32_t mode) {
  OFA_REQUIRE(mode == 0 || mode == 1, "ofa_train_side_mode: mode must be 0 or 1");
  const int old = g_side_mode;
  g_side_mode = mode;
  return old;
}

int64_t ofa_mbconv_train_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t cin, int32_t mid, int32_t cout) {
  (void)cin;
  if (n <= 0 || h <= 0 || w <= 0 || mid <= 0 || cout <= 0) return 0;
  return train_layout(n, h, w, mid, cout).total;
}

int ofa_mbconv_train_fwd(const OfaMBConvTrainArgs* a, void* stream) {
  int dev_n = 0;
  if (cudaGetDeviceCount(&dev_n) != cudaSuccess || dev_n == 0) {
    cudaGetLastError();
    return fail(OFA_ERR_CUDA, "no CUDA device: libofa_sr_b200 has no CPU fallback");
  }
  int rc = check_args(a, "ofa_mbconv_train_fwd");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const TrainLayout L = train_layout(a->n, a->h, a->w, a->mid, a->cout);
  char* ws = reinterpret_cast<char*>(a->ws);
  const int f16 = a->dtype == OFA_F16 ? 1 : 0;
  const OfaTensor4 x = nhwc16(a->x, a->dtype, a->n, a->cin, a->h, a->w);
  const OfaTensor4 y = nhwc16(a->y, a->dtype, a->n, a->cout, a->h, a->w);
  const OfaTensor4 z1 = nhwc16(ws + L.z1, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 a1 = nhwc16(ws + L.a1, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 z2 = nhwc16(ws + L.z2, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 a2 = nhwc16(ws + L.a2, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 z3 = nhwc16(ws + L.z3, a->dtype, a->n, a->cout, a->h, a->w);
  float* stats = reinterpret_cast<float*>(ws + L.stats);
  float *mean1 = stats, *var1 = stats + a->mid, *mean2 = stats + 2 * a->mid, *var2 = stats + 3 * a->mid,
        *mean3 = stats + 4 * a->mid, *var3 = stats + 4 * a->mid + a->cout;
  // the active weight slices, read in place from the full parameters: the two forward copies and the two transposed
  // copies the backward's data-gradient convs use (the weights cannot change in between: autograd's version check), one
  // launch for all four
  {
    OfaPackJob jobs[4];
    memset(jobs, 0, sizeof(jobs));
    const int dt = a->dtype;
    auto job = [&](int k, const float* w, int64_t so, int64_t si, int cin, int cout, void* out) {
      jobs[k].w = w; jobs[k].w_so = so; jobs[k].w_si = si; jobs[k].w_sh = 0; jobs[k].w_sw = 0;
      jobs[k].cin = cin; jobs[k].cout = cout; jobs[k].ks = 1; jobs[k].cin_pad = cin; jobs[k].cout_pad = (cout + 15) / 16 * 16;
      jobs[k].store = OFA_STORE_PLAIN; jobs[k].dtype = dt; jobs[k].out = out;
    };
    job(0, a->w_exp, a->w_exp_so, a->w_exp_si, a->cin, a->mid, ws + L.wexp);
    job(1, a->w_proj, a->w_proj_so, a->w_proj_si, a->mid, a->cout, ws + L.wproj);
    job(2, a->w_exp, a->w_exp_si, a->w_exp_so, a->mid, a->cin, ws + L.wexp_t);     // dX = conv(dZ1, W_exp^T): mid -> cin
    job(3, a->w_proj, a->w_proj_si, a->w_proj_so, a->cout, a->mid, ws + L.wproj_t); // dA2 = conv(dZ3, W_proj^T): cout -> mid
    if ((rc = launch_pack_weights4(jobs, 4, st))) return rc;
  }
  if ((rc = pointwise_tc(x, z1, ws + L.wexp, a->cin, a->mid, nullptr, st))) return rc;
  if ((rc = bn_fwd(z1, a1, a->bn_exp, mean1, var1, a->act, nullptr, st))) return rc;
  if (!dw_fast_supported(&a1, &z2, a->ks, nullptr))
    return fail(OFA_ERR_UNSUPPORTED, "mbconv training block: depthwise %d x %d on %d channels not supported", a->ks, a->ks, a->mid);
  if ((rc = launch_dw_fast(&a1, &z2, a->w_dw, a->kmax, a->m75, a->m53, a->transform_on, a->ks, 0, nullptr, st))) return rc;
  if ((rc = bn_fwd(z2, a2, a->bn_dw, mean2, var2, a->act, nullptr, st))) return rc;
  if ((rc = pointwise_tc(a2, z3, ws + L.wproj, a->mid, a->cout, nullptr, st))) return rc;
  return bn_fwd(z3, y, a->bn_proj, mean3, var3, OFA_ACT_NONE, a->add_residual ? &x : nullptr, st);
}

int ofa_mbconv_train_bwd(const OfaMBConvTrainArgs* a, const void* dy_ptr, void* dx_ptr, const OfaMBConvTrainGrads* g,
                         void* stream) {
  int dev_n = 0;
  if (cudaGetDeviceCount(&dev_n) != cudaSuccess || dev_n == 0) {
    cudaGetLastError();
    return fail(OFA_ERR_CUDA, "no CUDA device: libofa_sr_b200 has no CPU fallback");
  }
  int rc = check_args(a, "ofa_mbconv_train_bwd");
  if (rc) return rc;
  OFA_REQUIRE(dy_ptr && dx_ptr && g, "ofa_mbconv_train_bwd: null pointer");
  OFA_REQUIRE(((uintptr_t)dy_ptr & 15) == 0 && ((uintptr_t)dx_ptr & 15) == 0, "ofa_mbconv_train_bwd: dy / dx must be 16-byte aligned");
  OFA_REQUIRE(g->dw_exp && g->dw_dw && g->dw_proj, "ofa_mbconv_train_bwd: null weight gradient");
  for (int i = 0; i < 3; ++i) OFA_REQUIRE(g->dgamma[i] && g->dbeta[i], "ofa_mbconv_train_bwd: null BatchNorm gradient");
  if (a->transform_on && a->ks < a->kmax) {
    if (a->kmax == 7 && a->m75) OFA_REQUIRE(g->dm75 != nullptr, "ofa_mbconv_train_bwd: dm75 required");
    if (a->ks == 3) OFA_REQUIRE(g->dm53 != nullptr, "ofa_mbconv_train_bwd: dm53 required");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const TrainLayout L = train_layout(a->n, a->h, a->w, a->mid, a->cout);
  char* ws = reinterpret_cast<char*>(a->ws);
  const int f16 = a->dtype == OFA_F16 ? 1 : 0;
  const OfaTensor4 x = nhwc16(a->x, a->dtype, a->n, a->cin, a->h, a->w);
  const OfaTensor4 z1 = nhwc16(ws + L.z1, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 a1 = nhwc16(ws + L.a1, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 z2 = nhwc16(ws + L.z2, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 a2 = nhwc16(ws + L.a2, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 z3 = nhwc16(ws + L.z3, a->dtype, a->n, a->cout, a->h, a->w);
  const float* stats = reinterpret_cast<const float*>(ws + L.stats);
  const float *mean1 = stats, *var1 = stats + a->mid, *mean2 = stats + 2 * a->mid, *var2 = stats + 3 * a->mid,
              *mean3 = stats + 4 * a->mid, *var3 = stats + 4 * a->mid + a->cout;
  const OfaTensor4 dy = nhwc16(dy_ptr, a->dtype, a->n, a->cout, a->h, a->w);
  const OfaTensor4 dx = nhwc16(dx_ptr, a->dtype, a->n, a->cin, a->h, a->w);

  // stream-ordered scratch: dz3, four mid-wide gradient buffers (d(a2), d(z2), d(a1), d(z1): kept apart so the side
  // stream can still read a dZ while the main chain has moved on), the active-filter gradient
  const int64_t midb = align256(L.P * a->mid * 2), outb = align256(L.P * a->cout * 2);
  const int64_t dwab = align256((int64_t)sizeof(float) * a->mid * a->ks * a->ks);
  char* scratch = nullptr;
  keep_async_pool_resident();
  OFA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)(outb + 4 * midb + dwab), st));
  const OfaTensor4 dz3 = nhwc16(scratch, a->dtype, a->n, a->cout, a->h, a->w);
  const OfaTensor4 da2 = nhwc16(scratch + outb, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 dz2 = nhwc16(scratch + outb + midb, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 da1 = nhwc16(scratch + outb + 2 * midb, a->dtype, a->n, a->mid, a->h, a->w);
  const OfaTensor4 dz1 = nhwc16(scratch + outb + 3 * midb, a->dtype, a->n, a->mid, a->h, a->w);
  float* dwa = reinterpret_cast<float*>(scratch + outb + 4 * midb);
  const void* wproj_t = ws + L.wproj_t;
  const void* wexp_t = ws + L.wexp_t;
  SideCtx* side = side_ctx();
  cudaStream_t ss = side ? side->s : st;     // where the weight gradients run
  int forked = 0;
  auto fork = [&](int i) {                    // side stream continues from this point of the main stream
    if (!side) return;
    cudaEventRecord(side->fork[i], st);
    cudaStreamWaitEvent(side->s, side->fork[i], 0);
    forked = 1;
  };

  do {
    // BN3 (no activation; the residual branch passes dy through unchanged)
    if ((rc = bn_bwd(z3, dy, dz3, a->bn_proj, mean3, var3, OFA_ACT_NONE, g->dbeta[2], g->dgamma[2], st))) break;
    // project 1x1: weight gradient on tcgen05 (side), data gradient = conv of dz3 with W_proj^T (cout -> mid)
    if (!wgrad_tc_supported(&a2, &dz3, a->mid, a->cout, 1)) { rc = fail(OFA_ERR_UNSUPPORTED, "mbconv training block: project weight gradient"); break; }
    fork(0);
    if ((rc = launch_wgrad_tc(&a2, &dz3, g->dw_proj, a->w_proj_so, a->w_proj_si, 0, 0, a->mid, a->cout, 1, ss))) break;
    if ((rc = pointwise_tc(dz3, da2, wproj_t, a->cout, a->mid, nullptr, st))) break;
    // BN2 + act
    if ((rc = bn_bwd(z2, da2, dz2, a->bn_dw, mean2, var2, a->act, g->dbeta[1], g->dgamma[1], st))) break;
    // depthwise: filter gradient + chain rule through the 7->5->3 transforms (side), data gradient (rotated filter)
    fork(1);
    if ((rc = launch_dw_bwd_filter(make_tv(&a1), make_tv(&dz2), a->ks, dwa, ss))) break;
    if ((rc = launch_active_filter_bwd(a->w_dw, a->kmax, a->m75, a->m53, a->transform_on, a->ks, a->mid, dwa, g->dw_dw,
                                       g->dm75, g->dm53, ss))) break;
    if ((rc = launch_dw_fast(&dz2, &da1, a->w_dw, a->kmax, a->m75, a->m53, a->transform_on, a->ks, 1, nullptr, st))) break;
    // BN1 + act
    if ((rc = bn_bwd(z1, da1, dz1, a->bn_exp, mean1, var1, a->act, g->dbeta[0], g->dgamma[0], st))) break;
    // expand 1x1: weight gradient (side), dx = conv(dz1, W_exp^T) [+ dy: the identity branch]
    if (!wgrad_tc_supported(&x, &dz1, a->cin, a->mid, 1)) { rc = fail(OFA_ERR_UNSUPPORTED, "mbconv training block: expand weight gradient"); break; }
    fork(2);
    if ((rc = launch_wgrad_tc(&x, &dz1, g->dw_exp, a->w_exp_so, a->w_exp_si, 0, 0, a->cin, a->mid, 1, ss))) break;
    rc = pointwise_tc(dz1, dx, wexp_t, a->mid, a->cin, a->add_residual ? &dy : nullptr, st);
  } while (0);
  if (forked) {                               // join: everything the side stream was given is ordered before what follows
    cudaEventRecord(side->join, side->s);
    cudaStreamWaitEvent(st, side->join, 0);
  }
  cudaFreeAsync(scratch, st);
  return rc;
}

}  // extern "C"
