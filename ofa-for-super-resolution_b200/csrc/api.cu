// C ABI of libofa_sr_b200: argument checking and dispatch.  See include/ofa_sr_b200.h.
#include "ofa_common.cuh"
#include "kernels.h"

#include <stdlib.h>
#include <string.h>

namespace ofa {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
int64_t& launch_counter() {
  static thread_local int64_t n = 0;
  return n;
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(OFA_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  launch_counter()++;
  return OFA_OK;
}
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("OFA_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

static int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(OFA_ERR_CUDA, "no CUDA device: libofa_sr_b200 has no CPU fallback");
  }
  return OFA_OK;
}

static int check_transform_args(int kmax, const float* m75, const float* m53, int transform_on, int ks) {
  OFA_REQUIRE(kmax == 3 || kmax == 5 || kmax == 7, "kmax must be 3, 5 or 7 (got %d)", kmax);
  OFA_REQUIRE(ks == 3 || ks == 5 || ks == 7, "kernel size must be 3, 5 or 7 (got %d)", ks);
  OFA_REQUIRE(ks <= kmax, "active kernel size %d exceeds the stored %d", ks, kmax);
  if (transform_on && ks < kmax) {
    if (ks == 5) OFA_REQUIRE(m75 != nullptr, "7to5 matrix required for ks=5");
    if (ks == 3) OFA_REQUIRE(m53 != nullptr, "a ->3 transform matrix is required for ks=3");
  }
  return OFA_OK;
}

static int same_shape(const OfaTensor4* a, const OfaTensor4* b, const char* what) {
  OFA_REQUIRE(a->n == b->n && a->c == b->c && a->h == b->h && a->w == b->w,
              "%s: shape mismatch [%d,%d,%d,%d] vs [%d,%d,%d,%d]", what, a->n, a->c, a->h, a->w, b->n,
              b->c, b->h, b->w);
  return OFA_OK;
}

static int check_epi(const OfaEpilogue* e, const OfaTensor4* y) {
  if (!e) return OFA_OK;
  OFA_REQUIRE(e->act >= OFA_ACT_NONE && e->act <= OFA_ACT_RELU, "bad activation code %d", e->act);
  if (e->residual) {
    int rc = check_tensor(e->residual, "residual");
    if (rc) return rc;
    rc = same_shape(e->residual, y, "residual vs output");
    if (rc) return rc;
  }
  return OFA_OK;
}

static int store_shape_ok(const OfaConvArgs* a) {
  const OfaTensor4 &x = a->x, &y = a->y;
  if (a->store == OFA_STORE_PLAIN) {
    OFA_REQUIRE(y.n == x.n && y.c == a->cout && y.h == x.h && y.w == x.w, "conv: output shape mismatch");
  } else if (a->store == OFA_STORE_PIXELSHUFFLE2) {
    OFA_REQUIRE(a->cout % 4 == 0, "pixelshuffle needs cout %% 4 == 0");
    OFA_REQUIRE(y.n == x.n && y.c == a->cout / 4 && y.h == 2 * x.h && y.w == 2 * x.w,
                "conv+pixelshuffle: output shape mismatch");
  } else if (a->store == OFA_STORE_PIXELUNSHUFFLE2) {
    OFA_REQUIRE(x.h % 2 == 0 && x.w % 2 == 0, "pixelunshuffle needs even H and W");
    OFA_REQUIRE(y.n == x.n && y.c == a->cout * 4 && y.h == x.h / 2 && y.w == x.w / 2,
                "conv+pixelunshuffle: output shape mismatch");
  } else {
    return fail(OFA_ERR_ARG, "bad store mode %d", a->store);
  }
  return OFA_OK;
}

}  // namespace ofa

using namespace ofa;

extern "C" {

int ofa_version(void) { return 100; }
const char* ofa_last_error(void) { return err_buf(); }

int ofa_device_info(int32_t* sm, int32_t* major, int32_t* minor) {
  int rc = require_device();
  if (rc) return rc;
  int dev = 0;
  OFA_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  OFA_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  OFA_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  OFA_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm) *sm = a;
  if (major) *major = b;
  if (minor) *minor = c;
  return OFA_OK;
}

int64_t ofa_launch_count(void) { return launch_counter(); }
void ofa_launch_count_reset(void) { launch_counter() = 0; }

int ofa_dw_active_filter(const float* w7, int32_t kmax, const float* m75, const float* m53,
                         int32_t transform_on, int32_t ks, int32_t C, float* out, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(w7 && out, "ofa_dw_active_filter: null pointer");
  OFA_REQUIRE(C >= 0, "negative channel count");
  rc = check_transform_args(kmax, m75, m53, transform_on, ks);
  if (rc) return rc;
  return launch_active_filter(w7, kmax, m75, m53, transform_on, ks, C, out, (cudaStream_t)stream);
}

static int dw_common(const OfaTensor4* x, const OfaTensor4* y, const float* w7, int32_t kmax,
                     const float* m75, const float* m53, int32_t transform_on, int32_t ks, int flip,
                     const OfaEpilogue* epi, int32_t impl, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(y, "y"))) return rc;
  if ((rc = same_shape(x, y, "depthwise x vs y"))) return rc;
  OFA_REQUIRE(w7 != nullptr, "depthwise: null weight");
  if ((rc = check_transform_args(kmax, m75, m53, transform_on, ks))) return rc;
  if ((rc = check_epi(epi, y))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  bool fast_ok = dw_fast_supported(x, y, ks, epi);
  if (impl == OFA_IMPL_FAST && !fast_ok)
    return fail(OFA_ERR_UNSUPPORTED, "depthwise FAST path needs NHWC-dense bf16 x/y with C %% 64 == 0");
  if (fast_ok && impl != OFA_IMPL_SIMT)
    return launch_dw_fast(x, y, w7, kmax, m75, m53, transform_on, ks, flip, epi, st);
  return launch_dw_simt(make_tv(x), make_tv(y), w7, kmax, m75, m53, transform_on, ks, flip, make_epi(epi), st);
}

int ofa_dw_fwd(const OfaTensor4* x, const OfaTensor4* y, const float* w7, int32_t kmax, const float* m75,
               const float* m53, int32_t transform_on, int32_t ks, const OfaEpilogue* epi, int32_t impl,
               void* stream) {
  return dw_common(x, y, w7, kmax, m75, m53, transform_on, ks, 0, epi, impl, stream);
}

int ofa_dw_bwd_data(const OfaTensor4* dy, const OfaTensor4* dx, const float* w7, int32_t kmax,
                    const float* m75, const float* m53, int32_t transform_on, int32_t ks, void* stream) {
  return dw_common(dy, dx, w7, kmax, m75, m53, transform_on, ks, 1, nullptr, OFA_IMPL_AUTO, stream);
}

int ofa_conv_fwd(const OfaConvArgs* a, int32_t impl, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(a != nullptr, "ofa_conv_fwd: null args");
  if ((rc = check_tensor(&a->x, "x"))) return rc;
  if ((rc = check_tensor(&a->y, "y", /*allow_u8=*/true))) return rc;
  OFA_REQUIRE(a->ks >= 1 && (a->ks & 1), "conv kernel size must be odd (got %d)", a->ks);
  OFA_REQUIRE(a->cin >= 1 && a->cout >= 1, "conv: cin/cout must be positive");
  OFA_REQUIRE(a->x.c == a->cin, "conv: x has %d channels, cin = %d", a->x.c, a->cin);
  if ((rc = store_shape_ok(a))) return rc;
  if ((rc = check_epi(&a->epi, &a->y))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  OFA_REQUIRE(a->x.dtype != OFA_U8, "conv: uint8 is an OUTPUT format only");
  OFA_REQUIRE(a->y.dtype != OFA_U8 || (a->store == OFA_STORE_PLAIN && !a->epi.residual),
              "conv: uint8 output needs a plain store and no residual");
  if (impl != OFA_IMPL_SIMT && conv_out_rows_supported(a)) return launch_conv_out_rows(a, st);
  if (a->y.dtype == OFA_U8) {           // only the thin-output kernel and the CUDA-core kernel write the uint8 image
    OFA_REQUIRE(a->w != nullptr, "conv: null fp32 weight");
    return launch_conv_simt(make_tv(&a->x), make_tv(&a->y), a->w, a->w_so, a->w_si, a->w_sh, a->w_sw, a->cin,
                            a->cout, a->ks, a->flip, a->store, make_epi(&a->epi), st);
  }
  if (impl != OFA_IMPL_SIMT && conv_stem_tc_supported(a)) return launch_conv_stem_tc(a, st);
  if (impl != OFA_IMPL_SIMT && conv_stem_supported(a)) return launch_conv_stem(a, st);
  bool tc_ok = conv_tc_supported(a);
  if (impl == OFA_IMPL_FAST && !tc_ok)
    return fail(OFA_ERR_UNSUPPORTED, "conv FAST path: needs NHWC-dense bf16 x, packed bf16 weights, cin %% 64 == 0, cout <= 256 (or a multiple of 128/192)");
  if (tc_ok && impl != OFA_IMPL_SIMT) return launch_conv_tc(a, st);
  OFA_REQUIRE(a->w != nullptr, "conv: null fp32 weight");
  return launch_conv_simt(make_tv(&a->x), make_tv(&a->y), a->w, a->w_so, a->w_si, a->w_sh, a->w_sw, a->cin,
                          a->cout, a->ks, a->flip, a->store, make_epi(&a->epi), st);
}

int ofa_pw_fwd(const OfaConvArgs* a, int32_t impl, void* stream) {
  OFA_REQUIRE(a != nullptr && a->ks == 1, "ofa_pw_fwd requires ks == 1");
  return ofa_conv_fwd(a, impl, stream);
}
int ofa_conv_kxk_fwd(const OfaConvArgs* a, int32_t impl, void* stream) { return ofa_conv_fwd(a, impl, stream); }

int ofa_pack_weight_bf16(const float* w, int64_t w_so, int64_t w_si, int64_t w_sh, int64_t w_sw, int32_t cin,
                         int32_t cout, int32_t ks, int32_t cin_pad, int32_t cout_pad, int32_t store, void* out,
                         void* stream) {
  return ofa_pack_weight_16(w, w_so, w_si, w_sh, w_sw, cin, cout, ks, cin_pad, cout_pad, store, OFA_BF16, out, stream);
}

int ofa_pack_weight_16(const float* w, int64_t w_so, int64_t w_si, int64_t w_sh, int64_t w_sw, int32_t cin,
                       int32_t cout, int32_t ks, int32_t cin_pad, int32_t cout_pad, int32_t store, int32_t dtype,
                       void* out, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(w && out, "ofa_pack_weight_16: null pointer");
  OFA_REQUIRE(dtype == OFA_BF16 || dtype == OFA_F16, "ofa_pack_weight_16: dtype must be OFA_BF16 or OFA_F16");
  OFA_REQUIRE(cin_pad >= cin && cout_pad >= cout && cin >= 0 && cout >= 0 && ks >= 1, "bad pack dims");
  OFA_REQUIRE(store != OFA_STORE_PIXELSHUFFLE2 || cout % 4 == 0, "pixelshuffle pack needs cout %% 4 == 0");
  return launch_pack_weight(w, w_so, w_si, w_sh, w_sw, cin, cout, ks, cin_pad, cout_pad, store,
                            dtype == OFA_F16 ? 1 : 0, out, (cudaStream_t)stream);
}

int ofa_pack_weights_multi(const OfaPackJob* jobs_device, int32_t njobs, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(njobs >= 0 && njobs <= 65535, "ofa_pack_weights_multi: njobs out of range");
  OFA_REQUIRE(njobs == 0 || jobs_device, "ofa_pack_weights_multi: null job table");
  return launch_pack_weights_multi(jobs_device, njobs, (cudaStream_t)stream);
}

int ofa_bn_stats(const OfaTensor4* x, float* mean, float* var, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  OFA_REQUIRE(mean && var, "ofa_bn_stats: null output");
  OFA_REQUIRE((long long)x->n * x->h * x->w > 0, "ofa_bn_stats: empty batch");
  return launch_bn_stats(make_tv(x), mean, var, (cudaStream_t)stream);
}

int ofa_bn_update_running(const float* mean, const float* var, int64_t count, float* rm, float* rv,
                          float momentum, int32_t C, int64_t* num_batches_tracked, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(mean && var && rm && rv, "ofa_bn_update_running: null pointer");
  return launch_bn_update_running(mean, var, count, rm, rv, momentum, C,
                                  reinterpret_cast<long long*>(num_batches_tracked), (cudaStream_t)stream);
}

int ofa_affine_act(const OfaTensor4* x, const OfaTensor4* y, const OfaEpilogue* epi, int32_t store, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(y, "y"))) return rc;
  OfaConvArgs tmp;
  memset(&tmp, 0, sizeof(tmp));
  tmp.x = *x; tmp.y = *y; tmp.cout = x->c; tmp.store = store;
  if ((rc = store_shape_ok(&tmp))) return rc;
  if ((rc = check_epi(epi, y))) return rc;
  return launch_affine_act(make_tv(x), make_tv(y), make_epi(epi), store, (cudaStream_t)stream);
}

static int64_t mbconv_act_bytes(int32_t n, int32_t h, int32_t w, int32_t mid) {
  const int64_t P = (int64_t)n * h * w;
  const int64_t full = 2 * P * mid * 2;
  const int64_t nreg = (int64_t)n * ((h + 127) / 128) * ((w + 111) / 112);
  const int64_t band = mbconv_band_workspace_bytes(w, mid, (int)(nreg < (1 << 24) ? nreg : (1 << 24)));
  return (full > band ? full : band) + 256;
}

int64_t ofa_mbconv_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t cin, int32_t mid, int32_t cout) {
  (void)cin; (void)cout;
  int64_t P = (int64_t)n * h * w;
  int64_t mid_pad = (mid + 63) / 64 * 64;
  // two bf16 [P, mid] intermediates (or the band kernel's ring + counters, whichever is larger) + three packed bf16
  // weights (<= 384*64 each, padded) + slack
  return mbconv_act_bytes(n, h, w, (int32_t)mid_pad) + 2 * (int64_t)(384 + 64) * 384 * 2 + 4096;
}

int ofa_mbconv_fwd(const OfaMBConvArgs* a, int32_t impl, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(a != nullptr, "ofa_mbconv_fwd: null args");
  if ((rc = check_tensor(&a->x, "x"))) return rc;
  if ((rc = check_tensor(&a->y, "y"))) return rc;
  OFA_REQUIRE(is_nhwc_dense(&a->x) && is_nhwc_dense(&a->y) && is_16bit(a->x.dtype) && a->y.dtype == a->x.dtype,
              "ofa_mbconv_fwd: x and y must be NHWC-dense and both bf16 or both fp16");
  OFA_REQUIRE(a->x.c == a->cin && a->y.c == a->cout, "ofa_mbconv_fwd: channel mismatch");
  OFA_REQUIRE(a->cin % 64 == 0 && a->mid % 64 == 0 && a->cout % 64 == 0 && a->mid <= 384 && a->cin <= 384 && a->cout <= 256,
              "ofa_mbconv_fwd: cin/mid/cout must be multiples of 64 (cin,mid <= 384, cout <= 256)");
  OFA_REQUIRE(!a->add_residual || a->cin == a->cout, "ofa_mbconv_fwd: residual needs cin == cout");
  OFA_REQUIRE(a->w_exp && a->w_dw && a->w_proj, "ofa_mbconv_fwd: null weight");
  int64_t need = ofa_mbconv_workspace_bytes(a->x.n, a->x.h, a->x.w, a->cin, a->mid, a->cout);
  OFA_REQUIRE(a->ws && a->ws_bytes >= need, "ofa_mbconv_fwd: workspace too small (%lld < %lld)",
              (long long)a->ws_bytes, (long long)need);
  if ((rc = check_transform_args(a->kmax, a->m75, a->m53, a->transform_on, a->ks))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t P = (int64_t)a->x.n * a->x.h * a->x.w;
  char* ws = reinterpret_cast<char*>(a->ws);
  void* t1 = ws;
  void* t2 = ws + P * a->mid * 2;
  const int64_t act_bytes = mbconv_act_bytes(a->x.n, a->x.h, a->x.w, (a->mid + 63) / 64 * 64);
  void* wexp_p = ws + act_bytes;
  void* wproj_p = reinterpret_cast<char*>(wexp_p) + (int64_t)384 * 384 * 2;
  OFA_REQUIRE(a->mid_dtype == 0 || a->mid_dtype == OFA_BF16 || a->mid_dtype == OFA_F16, "bad mid_dtype %d",
              a->mid_dtype);

  // ---- planar tcgen05 path: expand -> Toeplitz depthwise -> project around channel-planar intermediates
  const bool force_planar = impl == OFA_IMPL_FAST || impl == OFA_IMPL_BAND || impl == OFA_IMPL_PLANAR3;
  if (impl != OFA_IMPL_SIMT && impl != OFA_IMPL_NHWC && mbconv_planar_supported(a) &&
      (force_planar || mbconv_planar_preferred(a))) {
    const int f16 = (a->mid_dtype == OFA_BF16) ? 0 : 1;
    const int tf16 = a->x.dtype == OFA_F16 ? 1 : 0;
    const int HW = a->x.h * a->x.w;
    const int mid_pad = (a->mid + 127) / 128 * 128;
    if (a->w_exp_packed && a->w_proj_packed) {            // the caller keeps a packed copy (valid until its weights change)
      wexp_p = const_cast<void*>(a->w_exp_packed);
      wproj_p = const_cast<void*>(a->w_proj_packed);
    } else if ((rc = launch_pack_block_weights(a->w_exp, a->w_exp_so, a->w_exp_si, a->w_proj, a->w_proj_so, a->w_proj_si,
                                               a->cin, a->mid, a->cout, mid_pad, tf16, f16, wexp_p, wproj_p, st))) {
      return rc;
    }
    if (impl != OFA_IMPL_PLANAR3 && mbconv_band_supported(a) && (impl == OFA_IMPL_BAND || mbconv_band_preferred(a)))
      return launch_mbconv_band(a, wexp_p, wproj_p, f16, ws, act_bytes, st);
    if ((rc = launch_expand_planar(a->x.ptr, t1, wexp_p, a->x.n, HW, a->mid, tf16, f16, &a->bn_exp, a->act, st)))
      return rc;
    if ((rc = launch_dw_planar(t1, t2, a->x.n, a->mid, a->x.h, a->x.w, a->w_dw, a->kmax, a->m75, a->m53,
                               a->transform_on, a->ks, f16, &a->bn_dw, a->act, st))) return rc;
    return launch_project_planar(t2, a->add_residual ? a->x.ptr : nullptr, a->y.ptr, wproj_p, a->x.n, HW, a->mid, tf16,
                                 f16, &a->bn_proj, st);
  }
  if (impl == OFA_IMPL_NHWC || force_planar) impl = impl == OFA_IMPL_FAST ? OFA_IMPL_FAST : OFA_IMPL_AUTO;
  // pack the active weight slices (tiny) — the slice W[:mid,:cin] is read in place from the full parameter
  const int xf16 = a->x.dtype == OFA_F16 ? 1 : 0;
  if ((rc = launch_pack_weight(a->w_exp, a->w_exp_so, a->w_exp_si, 0, 0, a->cin, a->mid, 1, a->cin, a->mid, OFA_STORE_PLAIN, xf16, wexp_p, st))) return rc;
  if ((rc = launch_pack_weight(a->w_proj, a->w_proj_so, a->w_proj_si, 0, 0, a->mid, a->cout, 1, a->mid, a->cout, OFA_STORE_PLAIN, xf16, wproj_p, st))) return rc;

  OfaTensor4 mid1 = a->x, mid2 = a->x;
  mid1.ptr = t1; mid2.ptr = t2;
  mid1.c = mid2.c = a->mid;
  mid1.sc = mid2.sc = 1;
  mid1.sw = mid2.sw = a->mid;
  mid1.sh = mid2.sh = (int64_t)a->x.w * a->mid;
  mid1.sn = mid2.sn = (int64_t)a->x.h * a->x.w * a->mid;
  mid1.dtype = mid2.dtype = a->x.dtype;

  OfaConvArgs c1;
  memset(&c1, 0, sizeof(c1));
  c1.x = a->x; c1.y = mid1; c1.w = a->w_exp; c1.w_so = a->w_exp_so; c1.w_si = a->w_exp_si;
  c1.w_bf16 = wexp_p; c1.cin_pad = a->cin; c1.cout_pad = a->mid; c1.cin = a->cin; c1.cout = a->mid; c1.ks = 1;
  c1.epi.gamma = a->bn_exp.gamma; c1.epi.beta = a->bn_exp.beta; c1.epi.mean = a->bn_exp.mean;
  c1.epi.var = a->bn_exp.var; c1.epi.eps = a->bn_exp.eps; c1.epi.act = a->act;
  if ((rc = ofa_conv_fwd(&c1, impl, stream))) return rc;

  OfaEpilogue e2;
  memset(&e2, 0, sizeof(e2));
  e2.gamma = a->bn_dw.gamma; e2.beta = a->bn_dw.beta; e2.mean = a->bn_dw.mean; e2.var = a->bn_dw.var;
  e2.eps = a->bn_dw.eps; e2.act = a->act;
  if ((rc = ofa_dw_fwd(&mid1, &mid2, a->w_dw, a->kmax, a->m75, a->m53, a->transform_on, a->ks, &e2, impl, stream))) return rc;

  OfaConvArgs c3;
  memset(&c3, 0, sizeof(c3));
  c3.x = mid2; c3.y = a->y; c3.w = a->w_proj; c3.w_so = a->w_proj_so; c3.w_si = a->w_proj_si;
  c3.w_bf16 = wproj_p; c3.cin_pad = a->mid; c3.cout_pad = a->cout; c3.cin = a->mid; c3.cout = a->cout; c3.ks = 1;
  c3.epi.gamma = a->bn_proj.gamma; c3.epi.beta = a->bn_proj.beta; c3.epi.mean = a->bn_proj.mean;
  c3.epi.var = a->bn_proj.var; c3.epi.eps = a->bn_proj.eps; c3.epi.act = OFA_ACT_NONE;
  c3.epi.residual = a->add_residual ? &a->x : nullptr;
  return ofa_conv_fwd(&c3, impl, stream);
}

static int dtype16_flag(int32_t dtype, int* f16) {
  OFA_REQUIRE(dtype == OFA_BF16 || dtype == OFA_F16, "planar stages store OFA_BF16 or OFA_F16 (got %d)", dtype);
  *f16 = dtype == OFA_F16 ? 1 : 0;
  return OFA_OK;
}

int ofa_mbconv_pack_weights(const float* w_exp, int64_t e_so, int64_t e_si, const float* w_proj, int64_t p_so,
                            int64_t p_si, int32_t mid, int32_t trunk_dtype, int32_t dtype, void* wexp_packed,
                            void* wproj_packed, void* stream) {
  int rc = require_device(), f16 = 0, tf16 = 0;
  if (rc) return rc;
  if ((rc = dtype16_flag(dtype, &f16))) return rc;
  if ((rc = dtype16_flag(trunk_dtype, &tf16))) return rc;
  OFA_REQUIRE(w_exp && w_proj && wexp_packed && wproj_packed, "ofa_mbconv_pack_weights: null pointer");
  OFA_REQUIRE(mid % 64 == 0 && mid >= 64 && mid <= 384, "mid must be a multiple of 64 in [64, 384]");
  return launch_pack_block_weights(w_exp, e_so, e_si, w_proj, p_so, p_si, 64, mid, 64, (mid + 127) / 128 * 128, tf16,
                                   f16, wexp_packed, wproj_packed, (cudaStream_t)stream);
}

static int planar_dims_ok(int32_t n, int64_t hw, int32_t mid) {
  OFA_REQUIRE(n > 0 && hw > 0 && hw % 8 == 0 && hw < (1ll << 31), "planar stages need N > 0 and H*W %% 8 == 0");
  OFA_REQUIRE(mid % 64 == 0 && mid >= 64 && mid <= 384, "mid must be a multiple of 64 in [64, 384]");
  return OFA_OK;
}

int ofa_expand_planar_fwd(const void* x_nhwc, void* y_planar, const void* wexp_packed, int32_t n, int32_t hw,
                          int32_t mid, int32_t trunk_dtype, int32_t dtype, const OfaBn* bn, int32_t act,
                          void* stream) {
  int rc = require_device(), f16 = 0, tf16 = 0;
  if (rc) return rc;
  if ((rc = dtype16_flag(dtype, &f16))) return rc;
  if ((rc = dtype16_flag(trunk_dtype, &tf16))) return rc;
  if ((rc = planar_dims_ok(n, hw, mid))) return rc;
  OFA_REQUIRE(x_nhwc && y_planar && wexp_packed && bn, "ofa_expand_planar_fwd: null pointer");
  OFA_REQUIRE(((uintptr_t)x_nhwc & 15) == 0 && ((uintptr_t)y_planar & 15) == 0 && ((uintptr_t)wexp_packed & 15) == 0,
              "ofa_expand_planar_fwd: pointers must be 16-byte aligned");
  return launch_expand_planar(x_nhwc, y_planar, wexp_packed, n, hw, mid, tf16, f16, bn, act, (cudaStream_t)stream);
}

int ofa_dw_planar_fwd(const void* x_planar, void* y_planar, int32_t n, int32_t c, int32_t h, int32_t w,
                      const float* w7, int32_t kmax, const float* m75, const float* m53, int32_t transform_on,
                      int32_t ks, int32_t dtype, const OfaBn* bn, int32_t act, void* stream) {
  int rc = require_device(), f16 = 0;
  if (rc) return rc;
  if ((rc = dtype16_flag(dtype, &f16))) return rc;
  OFA_REQUIRE(x_planar && y_planar && w7, "ofa_dw_planar_fwd: null pointer");
  OFA_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && w % 8 == 0, "ofa_dw_planar_fwd: needs W %% 8 == 0");
  OFA_REQUIRE(((uintptr_t)x_planar & 15) == 0 && ((uintptr_t)y_planar & 15) == 0, "pointers must be 16-byte aligned");
  if ((rc = check_transform_args(kmax, m75, m53, transform_on, ks))) return rc;
  return launch_dw_planar(x_planar, y_planar, n, c, h, w, w7, kmax, m75, m53, transform_on, ks, f16, bn, act,
                          (cudaStream_t)stream);
}

int ofa_project_planar_fwd(const void* x_planar, const void* res_nhwc, void* y_nhwc, const void* wproj_packed,
                           int32_t n, int32_t hw, int32_t mid, int32_t trunk_dtype, int32_t dtype, const OfaBn* bn,
                           void* stream) {
  int rc = require_device(), f16 = 0, tf16 = 0;
  if (rc) return rc;
  if ((rc = dtype16_flag(dtype, &f16))) return rc;
  if ((rc = dtype16_flag(trunk_dtype, &tf16))) return rc;
  if ((rc = planar_dims_ok(n, hw, mid))) return rc;
  OFA_REQUIRE(x_planar && y_nhwc && wproj_packed && bn, "ofa_project_planar_fwd: null pointer");
  OFA_REQUIRE(((uintptr_t)x_planar & 15) == 0 && ((uintptr_t)y_nhwc & 15) == 0 && ((uintptr_t)res_nhwc & 15) == 0 &&
                  ((uintptr_t)wproj_packed & 15) == 0, "pointers must be 16-byte aligned");
  return launch_project_planar(x_planar, res_nhwc, y_nhwc, wproj_packed, n, hw, mid, tf16, f16, bn,
                               (cudaStream_t)stream);
}

int ofa_adam_step(const OfaAdamTensor* table_dev, const int32_t* chunks_dev, int32_t n_tensors, int32_t n_chunks,
                  const float* const* grads_dev, int32_t* steps_dev, float lr, float beta1, float beta2, float eps,
                  void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(n_tensors >= 0 && n_chunks >= 0, "ofa_adam_step: negative counts");
  if (n_tensors == 0 || n_chunks == 0) return OFA_OK;
  OFA_REQUIRE(table_dev && chunks_dev && grads_dev && steps_dev, "ofa_adam_step: null pointer");
  OFA_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "ofa_adam_step: bad hyper-parameters");
  return launch_adam_step(table_dev, chunks_dev, n_tensors, n_chunks, grads_dev, steps_dev, lr, beta1, beta2, eps,
                          (cudaStream_t)stream);
}

int ofa_psnr_y_sse(const OfaTensor4* a, const OfaTensor4* b, int64_t* sse_per_image, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if (a && a->n == 0) return OFA_OK;
  if ((rc = check_tensor(a, "a"))) return rc;
  if ((rc = check_tensor(b, "b"))) return rc;
  if ((rc = same_shape(a, b, "psnr_y a vs b"))) return rc;
  OFA_REQUIRE(a->c == 3, "ofa_psnr_y_sse: RGB images expected (got %d channels)", a->c);
  OFA_REQUIRE(sse_per_image != nullptr, "ofa_psnr_y_sse: null output");
  OFA_REQUIRE((long long)a->h * a->w < (1ll << 31), "ofa_psnr_y_sse: image too large");
  return launch_psnr_y_sse(make_tv(a), make_tv(b), reinterpret_cast<long long*>(sse_per_image), (cudaStream_t)stream);
}

int ofa_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int32_t n,
                   int32_t in_features, int32_t out_features, int32_t act, float* y, int64_t ldy, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(n >= 0 && in_features >= 0 && out_features >= 0, "ofa_linear_fwd: negative size");
  if (n == 0 || out_features == 0) return OFA_OK;
  OFA_REQUIRE(x && w && y, "ofa_linear_fwd: null pointer");
  OFA_REQUIRE(ldx >= in_features && ldw >= in_features && ldy >= out_features, "ofa_linear_fwd: leading dimension too small");
  OFA_REQUIRE(act >= OFA_ACT_NONE && act <= OFA_ACT_HSIGMOID, "bad activation code %d", act);
  return launch_linear_fwd(x, ldx, w, ldw, bias, n, in_features, out_features, act, y, ldy, (cudaStream_t)stream);
}

int ofa_act_bwd_from_output(const float* dy, const float* y, int32_t act, float* dz, int64_t total, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if (total == 0) return OFA_OK;
  OFA_REQUIRE(dy && y && dz && total > 0, "ofa_act_bwd_from_output: bad arguments");
  OFA_REQUIRE(act == OFA_ACT_NONE || act == OFA_ACT_RELU || act == OFA_ACT_RELU6 || act == OFA_ACT_HSIGMOID,
              "ofa_act_bwd_from_output: activation %d has no derivative in terms of its output", act);
  return launch_act_bwd_from_output(dy, y, act, dz, total, (cudaStream_t)stream);
}

int ofa_linear_bwd_data(const float* dz, int64_t lddz, const float* w, int64_t ldw, int32_t n, int32_t in_features,
                        int32_t out_features, float* dx, int64_t lddx, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if (n == 0 || in_features == 0) return OFA_OK;
  OFA_REQUIRE(dz && w && dx && n > 0 && in_features > 0 && out_features >= 0, "ofa_linear_bwd_data: bad arguments");
  return launch_linear_bwd_data(dz, lddz, w, ldw, n, in_features, out_features, dx, lddx, (cudaStream_t)stream);
}

int ofa_linear_bwd_weight(const float* dz, int64_t lddz, const float* x, int64_t ldx, int32_t n, int32_t in_features,
                          int32_t out_features, float* dw, int64_t lddw, float* db, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if (out_features == 0) return OFA_OK;
  OFA_REQUIRE(dz && x && dw && n >= 0 && in_features >= 0 && out_features > 0, "ofa_linear_bwd_weight: bad arguments");
  return launch_linear_bwd_weight(dz, lddz, x, ldx, n, in_features, out_features, dw, lddw, db, (cudaStream_t)stream);
}

int ofa_plane_mean(const OfaTensor4* x, float* pooled, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  OFA_REQUIRE(pooled != nullptr, "ofa_plane_mean: null output");
  OFA_REQUIRE((long long)x->h * x->w > 0 && (long long)x->h * x->w < (1ll << 31), "ofa_plane_mean: bad plane size");
  TV tx = make_tv(x);
  return launch_plane_reduce(tx, nullptr, 1.f / (float)((long long)x->h * x->w), pooled, (cudaStream_t)stream);
}

int ofa_plane_dot(const OfaTensor4* x, const OfaTensor4* dy, float* out, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(dy, "dy"))) return rc;
  if ((rc = same_shape(x, dy, "plane_dot x vs dy"))) return rc;
  OFA_REQUIRE(out != nullptr, "ofa_plane_dot: null output");
  TV tx = make_tv(x), tdy = make_tv(dy);
  return launch_plane_reduce(tx, &tdy, 1.f, out, (cudaStream_t)stream);
}

int ofa_channel_scale(const OfaTensor4* x, const OfaTensor4* y, const float* s, const float* add, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(y, "y"))) return rc;
  if ((rc = same_shape(x, y, "channel_scale x vs y"))) return rc;
  OFA_REQUIRE(s != nullptr, "ofa_channel_scale: null scale");
  return launch_channel_scale(make_tv(x), make_tv(y), s, add, (cudaStream_t)stream);
}

static int dw_strided_shapes(const OfaTensor4* big, const OfaTensor4* small, int32_t ks, int32_t stride) {
  OFA_REQUIRE(ks == 3 || ks == 5 || ks == 7, "kernel size must be 3, 5 or 7 (got %d)", ks);
  OFA_REQUIRE(stride >= 1 && stride <= 4, "depthwise stride must be 1..4 (got %d)", stride);
  OFA_REQUIRE(small->n == big->n && small->c == big->c && small->h == (big->h - 1) / stride + 1 &&
                  small->w == (big->w - 1) / stride + 1,
              "strided depthwise: output must be [%d, %d, %d, %d]", big->n, big->c, (big->h - 1) / stride + 1,
              (big->w - 1) / stride + 1);
  return OFA_OK;
}

int ofa_dw_strided_fwd(const OfaTensor4* x, const OfaTensor4* y, const float* filt, int32_t ks, int32_t stride,
                       const OfaEpilogue* epi, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(y, "y"))) return rc;
  OFA_REQUIRE(filt != nullptr, "ofa_dw_strided_fwd: null filter");
  if ((rc = dw_strided_shapes(x, y, ks, stride))) return rc;
  if ((rc = check_epi(epi, y))) return rc;
  OFA_REQUIRE(!epi || !epi->residual, "ofa_dw_strided_fwd: no residual on a strided layer");
  return launch_dw_strided_fwd(make_tv(x), make_tv(y), filt, ks, stride, make_epi(epi), (cudaStream_t)stream);
}

int ofa_dw_strided_bwd_data(const OfaTensor4* dy, const OfaTensor4* dx, const float* filt, int32_t ks, int32_t stride,
                            void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(dy, "dy"))) return rc;
  if ((rc = check_tensor(dx, "dx"))) return rc;
  OFA_REQUIRE(filt != nullptr, "ofa_dw_strided_bwd_data: null filter");
  if ((rc = dw_strided_shapes(dx, dy, ks, stride))) return rc;
  return launch_dw_strided_bwd_data(make_tv(dy), make_tv(dx), filt, ks, stride, (cudaStream_t)stream);
}

int ofa_dw_strided_bwd_filter(const OfaTensor4* x, const OfaTensor4* dy, int32_t ks, int32_t stride, float* dfilt,
                              void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(dy, "dy"))) return rc;
  OFA_REQUIRE(dfilt != nullptr, "ofa_dw_strided_bwd_filter: null output");
  if ((rc = dw_strided_shapes(x, dy, ks, stride))) return rc;
  return launch_dw_strided_bwd_filter(make_tv(x), make_tv(dy), ks, stride, dfilt, (cudaStream_t)stream);
}

int32_t ofa_resample_ksize(int32_t in_size, int32_t out_size) { return resample_ksize(in_size, out_size); }

int ofa_resample_build_table(int32_t in_size, int32_t out_size, int32_t* bounds_host, int32_t* kk_host) {
  OFA_REQUIRE(in_size > 0 && out_size > 0, "ofa_resample_build_table: sizes must be positive");
  OFA_REQUIRE(bounds_host && kk_host, "ofa_resample_build_table: null output");
  return resample_build_table(in_size, out_size, bounds_host, kk_host);
}

int ofa_bicubic_resize_u8(const uint8_t* src, int32_t n, int32_t h, int32_t w, int32_t out_h, int32_t out_w,
                          const int32_t* bounds_h, const int32_t* kk_h, int32_t ksize_h, const int32_t* bounds_v,
                          const int32_t* kk_v, int32_t ksize_v, uint8_t* tmp, uint8_t* out_u8, float* out_f32,
                          void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(n >= 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0, "ofa_bicubic_resize_u8: bad sizes");
  if (n == 0) return OFA_OK;
  OFA_REQUIRE(src && tmp && (out_u8 || out_f32), "ofa_bicubic_resize_u8: null pointer");
  OFA_REQUIRE(bounds_h && kk_h && bounds_v && kk_v, "ofa_bicubic_resize_u8: null coefficient table");
  OFA_REQUIRE(ksize_h == resample_ksize(w, out_w) && ksize_v == resample_ksize(h, out_h),
              "ofa_bicubic_resize_u8: table built for other sizes (ksize %d / %d)", ksize_h, ksize_v);
  return launch_bicubic_resize_u8(src, n, h, w, out_h, out_w, bounds_h, kk_h, ksize_h, bounds_v, kk_v, ksize_v, tmp,
                                  out_u8, out_f32, (cudaStream_t)stream);
}

int ofa_sr_augment_u8(const uint8_t* src, int64_t sample_stride, int32_t n, int32_t h, int32_t w,
                      const int32_t* params, int32_t size, uint8_t* out_u8, float* out_f32, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(n >= 0 && h > 0 && w > 0 && size > 0 && size <= h && size <= w, "ofa_sr_augment_u8: bad sizes");
  if (n == 0) return OFA_OK;
  OFA_REQUIRE(src && params && (out_u8 || out_f32), "ofa_sr_augment_u8: null pointer");
  return launch_sr_augment_u8(src, sample_stride, n, w, params, size, out_u8, out_f32, (cudaStream_t)stream);
}

int ofa_dw_bwd_filter(const OfaTensor4* x, const OfaTensor4* dy, int32_t ks, float* dw_active, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(dy, "dy"))) return rc;
  if ((rc = same_shape(x, dy, "dw_bwd_filter x vs dy"))) return rc;
  OFA_REQUIRE(dw_active != nullptr, "null dw_active");
  return launch_dw_bwd_filter(make_tv(x), make_tv(dy), ks, dw_active, (cudaStream_t)stream);
}

int ofa_dw_active_filter_bwd(const float* w7, int32_t kmax, const float* m75, const float* m53,
                             int32_t transform_on, int32_t ks, int32_t C, const float* dw_active, float* dw7,
                             float* dm75, float* dm53, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  OFA_REQUIRE(w7 && dw_active && dw7, "ofa_dw_active_filter_bwd: null pointer");
  if ((rc = check_transform_args(kmax, m75, m53, transform_on, ks))) return rc;
  if (transform_on && ks < kmax) {
    if (kmax == 7 && m75) OFA_REQUIRE(dm75 != nullptr, "dm75 required");
    if (ks == 3) OFA_REQUIRE(dm53 != nullptr, "dm53 required");
  }
  return launch_active_filter_bwd(w7, kmax, m75, m53, transform_on, ks, C, dw_active, dw7, dm75, dm53,
                                  (cudaStream_t)stream);
}

int ofa_conv_bwd_data(const OfaTensor4* dy, const OfaTensor4* dx, const float* w, int64_t w_so, int64_t w_si,
                      int64_t w_sh, int64_t w_sw, int32_t cin, int32_t cout, int32_t ks, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(dy, "dy"))) return rc;
  if ((rc = check_tensor(dx, "dx"))) return rc;
  OFA_REQUIRE(w != nullptr, "null weight");
  OFA_REQUIRE(dy->c == cout && dx->c == cin && dy->n == dx->n && dy->h == dx->h && dy->w == dx->w,
              "conv_bwd_data: shape mismatch");
  // dX[p, i] = sum_{tap, o} dY[p - tap, o] W[o, i, tap]  ==  conv of dY with (o <-> i swapped, flipped) W
  Epi none = make_epi(nullptr);
  return launch_conv_simt(make_tv(dy), make_tv(dx), w, /*so=*/w_si, /*si=*/w_so, w_sh, w_sw, /*cin=*/cout,
                          /*cout=*/cin, ks, /*flip=*/1, OFA_STORE_PLAIN, none, (cudaStream_t)stream);
}

int ofa_conv_bwd_weight(const OfaTensor4* x, const OfaTensor4* dy, float* dw, int64_t w_so, int64_t w_si,
                        int64_t w_sh, int64_t w_sw, int32_t cin, int32_t cout, int32_t ks, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(dy, "dy"))) return rc;
  OFA_REQUIRE(dw != nullptr, "null dw");
  OFA_REQUIRE(x->c == cin && dy->c == cout && dy->n == x->n && dy->h == x->h && dy->w == x->w,
              "conv_bwd_weight: shape mismatch");
  if (wgrad_tc_supported(x, dy, cin, cout, ks))
    return launch_wgrad_tc(x, dy, dw, w_so, w_si, w_sh, w_sw, cin, cout, ks, (cudaStream_t)stream);
  return launch_conv_bwd_weight(make_tv(x), make_tv(dy), dw, w_so, w_si, w_sh, w_sw, cin, cout, ks,
                                (cudaStream_t)stream);
}

int ofa_bn_train_fwd(const OfaTensor4* x, const OfaTensor4* y, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int32_t act, const OfaTensor4* residual,
                     float* batch_mean, float* batch_var, int64_t* num_batches_tracked, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(y, "y"))) return rc;
  if ((rc = same_shape(x, y, "bn_train_fwd x vs y"))) return rc;
  OFA_REQUIRE(batch_mean && batch_var, "ofa_bn_train_fwd: null statistics output");
  OFA_REQUIRE((long long)x->n * x->h * x->w > 0, "ofa_bn_train_fwd: empty batch");
  OFA_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "ofa_bn_train_fwd: running_mean / running_var");
  cudaStream_t st = (cudaStream_t)stream;
  const bool upd = running_mean != nullptr && momentum != 0.f;
  if ((rc = launch_bn_stats_update(make_tv(x), batch_mean, batch_var, upd ? running_mean : nullptr,
                                   upd ? running_var : nullptr, momentum,
                                   upd ? reinterpret_cast<long long*>(num_batches_tracked) : nullptr, st))) return rc;
  OfaEpilogue e;
  memset(&e, 0, sizeof(e));
  e.gamma = gamma; e.beta = beta; e.mean = batch_mean; e.var = batch_var; e.eps = eps; e.act = act; e.residual = residual;
  if ((rc = check_epi(&e, y))) return rc;
  return launch_affine_act(make_tv(x), make_tv(y), make_epi(&e), OFA_STORE_PLAIN, st);
}

int ofa_bn_train_bwd(const OfaTensor4* x, const OfaTensor4* dy, const OfaTensor4* dx, const float* gamma,
                     const float* beta, const float* mean, const float* var, float eps, int32_t act, int32_t training,
                     float* sum_dz, float* sum_dz_xhat, void* stream) {
  int rc = ofa_bn_bwd_reduce(x, dy, gamma, beta, mean, var, eps, act, sum_dz, sum_dz_xhat, stream);
  if (rc || dx == nullptr) return rc;
  return ofa_bn_bwd_apply(x, dy, dx, gamma, beta, mean, var, eps, act, training, sum_dz, sum_dz_xhat, stream);
}

int ofa_bn_bwd_reduce(const OfaTensor4* x, const OfaTensor4* dy, const float* gamma, const float* beta,
                      const float* mean, const float* var, float eps, int32_t act, float* sum_dz,
                      float* sum_dz_xhat, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(dy, "dy"))) return rc;
  if ((rc = same_shape(x, dy, "bn_bwd_reduce x vs dy"))) return rc;
  OFA_REQUIRE(mean && var && sum_dz && sum_dz_xhat, "ofa_bn_bwd_reduce: null pointer");
  return launch_bn_bwd_reduce(make_tv(x), make_tv(dy), gamma, beta, mean, var, eps, act, sum_dz, sum_dz_xhat,
                              (cudaStream_t)stream);
}

int ofa_bn_bwd_apply(const OfaTensor4* x, const OfaTensor4* dy, const OfaTensor4* dx, const float* gamma,
                     const float* beta, const float* mean, const float* var, float eps, int32_t act,
                     int32_t training, const float* sum_dz, const float* sum_dz_xhat, void* stream) {
  int rc = require_device();
  if (rc) return rc;
  if ((rc = check_tensor(x, "x"))) return rc;
  if ((rc = check_tensor(dy, "dy"))) return rc;
  if ((rc = check_tensor(dx, "dx"))) return rc;
  if ((rc = same_shape(x, dy, "bn_bwd_apply x vs dy"))) return rc;
  if ((rc = same_shape(x, dx, "bn_bwd_apply x vs dx"))) return rc;
  if (training) OFA_REQUIRE(sum_dz && sum_dz_xhat && mean && var, "bn_bwd_apply(training): null stats");
  return launch_bn_bwd_apply(make_tv(x), make_tv(dy), make_tv(dx), gamma, beta, mean, var, eps, act, training,
                             sum_dz, sum_dz_xhat, (cudaStream_t)stream);
}

}  // extern "C"
