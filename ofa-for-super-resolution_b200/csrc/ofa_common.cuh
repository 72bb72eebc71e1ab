// Shared host/device helpers of libofa_sr_b200 (B200 / sm_100a only).
#pragma once
#include <string.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/ofa_sr_b200.h"

namespace ofa {

// ------------------------------------------------------------------------------------------------
// per-thread error text + launch counter (the only mutable state of the library)
// ------------------------------------------------------------------------------------------------
char* err_buf();
int64_t& launch_counter();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError after a launch; bumps the launch counter
int sm_count();                      // cached per device

#define OFA_REQUIRE(cond, ...)                        \
  do {                                                \
    if (!(cond)) return ::ofa::fail(OFA_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define OFA_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ::ofa::fail(OFA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  The training step is ~550 short launches (a few to a few tens of microseconds each)
// in one stream; with plain stream ordering every kernel's launch latency and block ramp-up sit between two kernels.
// Kernels that start with pdl_wait() may be launched with launch_pdl(): their blocks are scheduled while the
// preceding kernel drains and block in griddepcontrol.wait until it has completed and its writes are visible --
// the same ordering as before, minus the launch gap.  pdl_wait() must come before the first global-memory access;
// without the launch attribute it is a no-op.  OFA_PDL=0 launches everything with plain stream ordering.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

bool pdl_enabled();   // api.cu: OFA_PDL != "0"

// true the first time it is called with `done` on the current device: per-function launch attributes (dynamic
// shared-memory size) are set once per device instead of before every launch (a driver call per launch)
inline bool once_per_device(unsigned char (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = 1;
  return true;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------
// strided 4-D activation view usable on the device
// ------------------------------------------------------------------------------------------------
struct TV {
  void* ptr;
  int dtype;
  int n, c, h, w;
  long long sn, sc, sh, sw;

  __host__ __device__ __forceinline__ long long off(int in_, int ic, int ih, int iw) const {
    return in_ * sn + ic * sc + ih * sh + iw * sw;
  }
  __device__ __forceinline__ float ld(long long o) const {
    if (dtype == OFA_F32) return reinterpret_cast<const float*>(ptr)[o];
    if (dtype == OFA_F16) return __half2float(reinterpret_cast<const __half*>(ptr)[o]);
    if (dtype == OFA_U8) return (float)reinterpret_cast<const uint8_t*>(ptr)[o] * (1.f / 255.f);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ptr)[o]);
  }
  __device__ __forceinline__ void st(long long o, float v) const {
    if (dtype == OFA_F32)
      reinterpret_cast<float*>(ptr)[o] = v;
    else if (dtype == OFA_F16)
      reinterpret_cast<__half*>(ptr)[o] = __float2half_rn(v);
    else if (dtype == OFA_U8)      // tensor2img_np: clamp_(0, 1) -> * 255.0 -> numpy round (half to even) -> uint8
      reinterpret_cast<uint8_t*>(ptr)[o] = (uint8_t)rintf(fminf(fmaxf(v, 0.f), 1.f) * 255.f);
    else
      reinterpret_cast<__nv_bfloat16*>(ptr)[o] = __float2bfloat16_rn(v);
  }
};

inline TV make_tv(const OfaTensor4* t) {
  TV v;
  v.ptr = t->ptr;
  v.dtype = t->dtype;
  v.n = t->n; v.c = t->c; v.h = t->h; v.w = t->w;
  v.sn = t->sn; v.sc = t->sc; v.sh = t->sh; v.sw = t->sw;
  return v;
}
inline TV null_tv() {
  TV v;
  v.ptr = nullptr; v.dtype = 0; v.n = v.c = v.h = v.w = 0; v.sn = v.sc = v.sh = v.sw = 0;
  return v;
}
inline bool is_nhwc_dense(const OfaTensor4* t) {
  return t->sc == 1 && t->sw == t->c && t->sh == (int64_t)t->w * t->c &&
         t->sn == (int64_t)t->h * t->w * t->c;
}
inline bool c_inner(const OfaTensor4* t) { return t->sc == 1; }
inline int check_tensor(const OfaTensor4* t, const char* name, bool allow_u8 = false) {
  if (!t) return fail(OFA_ERR_ARG, "%s: null tensor", name);
  if (!t->ptr) return fail(OFA_ERR_ARG, "%s: null data pointer", name);
  if (t->dtype != OFA_F32 && t->dtype != OFA_BF16 && t->dtype != OFA_F16 && !(allow_u8 && t->dtype == OFA_U8))
    return fail(OFA_ERR_ARG, "%s: bad dtype %d", name, t->dtype);
  if (t->n < 0 || t->c < 0 || t->h < 0 || t->w < 0) return fail(OFA_ERR_ARG, "%s: negative extent", name);
  return OFA_OK;
}
inline long long numel(const OfaTensor4* t) { return (long long)t->n * t->c * t->h * t->w; }

inline bool is_16bit(int dtype) { return dtype == OFA_BF16 || dtype == OFA_F16; }

// two packed 16-bit floats (bf16 or fp16) <-> two fp32
__device__ __forceinline__ uint32_t pack16(float a, float b, int f16) {
  if (f16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// pack16(relu6(a), relu6(b)) in two instructions: the conversion clamps at zero (cvt.rn.relu), a packed min clamps at 6.
// Bit-identical to clamping in fp32 first: rounding is monotonic and 0 and 6 are exact in both 16-bit formats.
__device__ __forceinline__ uint32_t pack16_relu6(float a, float b, int f16) {
  uint32_t r;
  if (f16) {
    asm("{\n\t.reg .b32 t;\n\tcvt.rn.relu.f16x2.f32 t, %2, %1;\n\tmin.f16x2 %0, t, %3;\n\t}" : "=r"(r) : "f"(a), "f"(b), "r"(0x46004600u));
  } else {
    asm("{\n\t.reg .b32 t;\n\tcvt.rn.relu.bf16x2.f32 t, %2, %1;\n\tmin.bf16x2 %0, t, %3;\n\t}" : "=r"(r) : "f"(a), "f"(b), "r"(0x40C040C0u));
  }
  return r;
}
__device__ __forceinline__ float2 unpack16(uint32_t u, int f16) {
  if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint16_t cvt16(float a, int f16) {
  if (f16) {
    __half h = __float2half_rn(a);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __nv_bfloat16 h = __float2bfloat16_rn(a);
  return *reinterpret_cast<uint16_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// epilogue on the device: per-channel affine (BN fold) + activation + residual
// ------------------------------------------------------------------------------------------------
struct Epi {
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* var;
  float eps;
  int act;
  TV res;  // res.ptr == nullptr -> none
};

inline Epi make_epi(const OfaEpilogue* e) {
  Epi d;
  if (!e) {
    d.gamma = d.beta = d.mean = d.var = nullptr;
    d.eps = 0.f; d.act = OFA_ACT_NONE; d.res = null_tv();
    return d;
  }
  d.gamma = e->gamma; d.beta = e->beta; d.mean = e->mean; d.var = e->var;
  d.eps = e->eps; d.act = e->act;
  d.res = e->residual ? make_tv(e->residual) : null_tv();
  return d;
}

__device__ __forceinline__ void epi_scale_shift(const Epi& e, int c, float& scale, float& shift) {
  float g = e.gamma ? e.gamma[c] : 1.f;
  float b = e.beta ? e.beta[c] : 0.f;
  float m = e.mean ? e.mean[c] : 0.f;
  float rstd = e.var ? rsqrtf(e.var[c] + e.eps) : 1.f;
  scale = g * rstd;
  shift = b - m * scale;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case OFA_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
    case OFA_ACT_RELU: return fmaxf(v, 0.f);
    case OFA_ACT_HSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
    default: return v;
  }
}
// derivative of the activation w.r.t. its input z
__device__ __forceinline__ float act_grad(float z, int act) {
  switch (act) {
    case OFA_ACT_RELU6: return (z > 0.f && z < 6.f) ? 1.f : 0.f;
    case OFA_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case OFA_ACT_HSWISH:
      if (z <= -3.f) return 0.f;
      if (z >= 3.f) return 1.f;
      return (2.f * z + 3.f) * (1.f / 6.f);
    default: return 1.f;
  }
}

// conv-resolution coordinate (c, h, w) -> coordinate in the stored tensor for a store mode
__device__ __forceinline__ void store_coord(int store, int c, int h, int w, int& oc, int& oh, int& ow) {
  if (store == OFA_STORE_PIXELSHUFFLE2) {
    oc = c >> 2; oh = 2 * h + ((c >> 1) & 1); ow = 2 * w + (c & 1);
  } else if (store == OFA_STORE_PIXELUNSHUFFLE2) {
    oc = 4 * c + 2 * (h & 1) + (w & 1); oh = h >> 1; ow = w >> 1;
  } else {
    oc = c; oh = h; ow = w;
  }
}

// ------------------------------------------------------------------------------------------------
// elastic depthwise filter (dynamic_op.py:46-71) computed for one channel into `out[ks*ks]`
//   w: the channel's kmax*kmax weights.  tmp: >= 25 floats of scratch.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void active_filter_channel(const float* __restrict__ w, int kmax,
                                                      const float* __restrict__ m75,
                                                      const float* __restrict__ m53, int transform_on,
                                                      int ks, float* out, float* tmp) {
  if (!transform_on || ks == kmax) {
    const int s = kmax / 2 - ks / 2;
    for (int y = 0; y < ks; ++y)
      for (int x = 0; x < ks; ++x) out[y * ks + x] = w[(y + s) * kmax + (x + s)];
    return;
  }
  // chain of learned transforms, largest first (F.linear: out[j] = sum_i in[i] * M[j][i]).
  //   kmax 7, m75 given : centre 5x5 of the 7x7 through the 25x25 matrix
  //   then (ks == 3)    : centre 3x3 of the current filter through the 9x9 matrix m53
  //   (kmax 5 -> 3 and a direct 7 -> 3 matrix are the same second step with m75 == NULL)
  const float* cur = w;
  int kc = kmax;
  if (kmax == 7 && m75 != nullptr) {
    for (int j = 0; j < 25; ++j) {
      float acc = 0.f;
      for (int i = 0; i < 25; ++i) acc = fmaf(w[(i / 5 + 1) * 7 + (i % 5 + 1)], m75[j * 25 + i], acc);
      tmp[j] = acc;
    }
    cur = tmp;
    kc = 5;
  }
  if (ks == kc) {
    for (int j = 0; j < ks * ks; ++j) out[j] = cur[j];
    return;
  }
  const int s = kc / 2 - 1;  // ks == 3
  for (int j = 0; j < 9; ++j) {
    float acc = 0.f;
    for (int i = 0; i < 9; ++i) acc = fmaf(cur[(i / 3 + s) * kc + (i % 3 + s)], m53[j * 9 + i], acc);
    out[j] = acc;
  }
}

}  // namespace ofa
