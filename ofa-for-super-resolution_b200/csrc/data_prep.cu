// SR data preparation on the device: what the reference's CPU data-loader workers do per
// sample with Pillow / torchvision (ofa/imagenet_codebase/data_providers/div2k_setxx.py:166-171, 288-298, 355-380):
//   RandomCrop -> RandomHorizontalFlip -> RandomRotation on the uint8 HR image, Scale(1/2) and Scale(1/4) with
//   Image.BICUBIC, ToTensor on all three.
// Everything is integer / byte work and bit-exact against Pillow:
//   * resize = Pillow's two-pass separable resampling (libImaging/Resample.c): per output index a window
//     [xmin, xmin + count) and 22-bit fixed-point coefficients (the cubic's support is scaled by the down-scaling
//     factor, windows are clipped and re-normalised at the borders); a pass is  clip8((2^21 + sum px * k) >> 22)
//     and the intermediate between the passes is uint8.  The tables are built on the HOST in double precision
//     (ofa_resample_build_table, the same operation order as precompute_coeffs / normalize_coeffs_8bpc).
//   * rotation = the nearest-neighbour affine walk of libImaging/Geometry.c in 16.16 fixed point; the host hands
//     over the six fixed-point coefficients per sample.
//   * ToTensor = (float)u8 / 255.0f (IEEE division) into a planar fp32 image.
// HBM-bound byte kernels: thread per output pixel (3 channels), output-major indexing so stores are coalesced; the
// windows of neighbouring threads overlap and are served by L1.
#include "ofa_common.cuh"
#include "kernels.h"

#include <math.h>

namespace ofa {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

// ---- host: coefficient tables -----------------------------------------------------------------------
static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

int resample_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  double filterscale = (double)in_size / (double)out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  return (int)ceil(support) * 2 + 1;
}

int resample_build_table(int in_size, int out_size, int32_t* bounds, int32_t* kk) {
  const int ksize = resample_ksize(in_size, out_size);
  if (ksize <= 0) return fail(OFA_ERR_ARG, "resample table: sizes must be positive");
  const double scale = (double)in_size / (double)out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const double ss = 1.0 / filterscale;
  double* w = new double[ksize];
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    int32_t* k = kk + (size_t)xx * ksize;
    for (int x = 0; x < ksize; ++x) k[x] = 0;
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      k[x] = v < 0 ? (int32_t)(-0.5 + v * (1 << RS_PRECISION_BITS)) : (int32_t)(0.5 + v * (1 << RS_PRECISION_BITS));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  delete[] w;
  return OFA_OK;
}

// ---- device ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= RS_PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
}

// horizontal pass: src [N][H][W][3] -> tmp [N][H][ow][3]
__global__ void resample_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ tmp, int H, int W, int ow,
                                  const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                                  unsigned total) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned ox = i % (unsigned)ow, row = i / (unsigned)ow;          // row = n * H + y
    const int xmin = bounds[2 * ox], cnt = bounds[2 * ox + 1];
    const int32_t* k = kk + (size_t)ox * ksize;
    const uint8_t* p = src + ((size_t)row * W + xmin) * 3;
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < cnt; ++x) {
      const int c = k[x];
      s0 += p[3 * x] * c; s1 += p[3 * x + 1] * c; s2 += p[3 * x + 2] * c;
    }
    uint8_t* o = tmp + (size_t)i * 3;
    o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
  }
}

// vertical pass: tmp [N][H][ow][3] -> uint8 [N][oh][ow][3] and / or fp32 [N][3][oh][ow] = value / 255
__global__ void resample_v_kernel(const uint8_t* __restrict__ tmp, uint8_t* __restrict__ out_u8,
                                  float* __restrict__ out_f32, int H, int oh, int ow,
                                  const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                                  unsigned total) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned ox = i % (unsigned)ow;
    unsigned r = i / (unsigned)ow;
    const unsigned oy = r % (unsigned)oh, n = r / (unsigned)oh;
    const int ymin = bounds[2 * oy], cnt = bounds[2 * oy + 1];
    const int32_t* k = kk + (size_t)oy * ksize;
    const uint8_t* p = tmp + (((size_t)n * H + ymin) * ow + ox) * 3;
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < cnt; ++y) {
      const int c = k[y];
      const uint8_t* q = p + (size_t)y * ow * 3;
      s0 += q[0] * c; s1 += q[1] * c; s2 += q[2] * c;
    }
    const uint8_t v0 = clip8(s0), v1 = clip8(s1), v2 = clip8(s2);
    if (out_u8) { uint8_t* o = out_u8 + (size_t)i * 3; o[0] = v0; o[1] = v1; o[2] = v2; }
    if (out_f32) {
      const size_t plane = (size_t)oh * ow;
      float* o = out_f32 + (size_t)n * 3 * plane + (size_t)oy * ow + ox;
      o[0] = (float)v0 / 255.0f; o[plane] = (float)v1 / 255.0f; o[2 * plane] = (float)v2 / 255.0f;
    }
  }
}

int launch_bicubic_resize_u8(const uint8_t* src, int N, int H, int W, int oh, int ow, const int32_t* bounds_h,
                             const int32_t* kk_h, int ksize_h, const int32_t* bounds_v, const int32_t* kk_v,
                             int ksize_v, uint8_t* tmp, uint8_t* out_u8, float* out_f32, cudaStream_t st) {
  const long long t1 = (long long)N * H * ow, t2 = (long long)N * oh * ow;
  if (t1 == 0 || t2 == 0) return OFA_OK;
  if (t1 >= (1ll << 32) || t2 >= (1ll << 32)) return fail(OFA_ERR_UNSUPPORTED, "bicubic resize: batch too large");
  const long long cap = (long long)sm_count() * 8;
  long long b1 = (t1 + 255) / 256, b2 = (t2 + 255) / 256;
  if (b1 > cap) b1 = cap;
  if (b2 > cap) b2 = cap;
  resample_h_kernel<<<(unsigned)b1, 256, 0, st>>>(src, tmp, H, W, ow, bounds_h, kk_h, ksize_h, (unsigned)t1);
  int rc = check_launch("resample_h_kernel");
  if (rc) return rc;
  resample_v_kernel<<<(unsigned)b2, 256, 0, st>>>(tmp, out_u8, out_f32, H, oh, ow, bounds_v, kk_v, ksize_v,
                                                  (unsigned)t2);
  return check_launch("resample_v_kernel");
}

// crop -> horizontal flip -> rotation of one S x S patch per sample.
//   params[n] = { i, j, flip, mode, a0, a1, a2, a3, a4, a5 }: crop origin (row i, column j), flip flag, rotation
//   mode (0 none, 1 = 180 deg, 2 = 90 deg, 3 = 270 deg as Pillow's transposes, 4 = affine walk) and the 16.16
//   fixed-point inverse matrix of Geometry.c affine_fixed: xin = (a2 + a1 * y + a0 * x) >> 16, yin likewise with
//   a5, a4, a3; pixels that map outside the patch are 0.
__global__ void augment_kernel(const uint8_t* __restrict__ src, long long sample_stride, int W,
                               const int32_t* __restrict__ params, int S, uint8_t* __restrict__ out_u8,
                               float* __restrict__ out_f32, unsigned total) {
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int x = (int)(idx % (unsigned)S);
    unsigned r = idx / (unsigned)S;
    const int y = (int)(r % (unsigned)S);
    const unsigned n = r / (unsigned)S;
    const int32_t* q = params + (size_t)n * 10;
    int xin, yin;
    switch (q[3]) {
      case 0: xin = x; yin = y; break;
      case 1: xin = S - 1 - x; yin = S - 1 - y; break;
      case 2: xin = S - 1 - y; yin = x; break;           // counter-clockwise quarter turn: out[y][x] = in[x][S-1-y]
      case 3: xin = y; yin = S - 1 - x; break;
      default:
        xin = (int)(((long long)q[6] + (long long)q[5] * y + (long long)q[4] * x) >> 16);
        yin = (int)(((long long)q[9] + (long long)q[8] * y + (long long)q[7] * x) >> 16);
    }
    uint8_t v0 = 0, v1 = 0, v2 = 0;
    if (xin >= 0 && xin < S && yin >= 0 && yin < S) {
      const int sx = q[1] + (q[2] ? S - 1 - xin : xin), sy = q[0] + yin;
      const uint8_t* p = src + (size_t)n * sample_stride + ((size_t)sy * W + sx) * 3;
      v0 = p[0]; v1 = p[1]; v2 = p[2];
    }
    if (out_u8) { uint8_t* o = out_u8 + (size_t)idx * 3; o[0] = v0; o[1] = v1; o[2] = v2; }
    if (out_f32) {
      const size_t plane = (size_t)S * S;
      float* o = out_f32 + (size_t)n * 3 * plane + (size_t)y * S + x;
      o[0] = (float)v0 / 255.0f; o[plane] = (float)v1 / 255.0f; o[2 * plane] = (float)v2 / 255.0f;
    }
  }
}

int launch_sr_augment_u8(const uint8_t* src, long long sample_stride, int N, int W, const int32_t* params, int S,
                         uint8_t* out_u8, float* out_f32, cudaStream_t st) {
  const long long total = (long long)N * S * S;
  if (total == 0) return OFA_OK;
  if (total >= (1ll << 32)) return fail(OFA_ERR_UNSUPPORTED, "augment: batch too large");
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  augment_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, sample_stride, W, params, S, out_u8, out_f32, (unsigned)total);
  return check_launch("augment_kernel");
}

}  // namespace ofa
