// The MobileNetV3 flavour of the elastic modules: ops the SR nets do not exercise but the
// ofa/elastic_nn API exposes.  Exact fp32 CUDA-core kernels, any tensor layout / dtype through TV.
//   * squeeze-and-excite, DynamicSE (dynamic_op.py:175-200, ofa/utils.py:354-375): global average pool ->
//     sliced 1x1 reduce (+bias, ReLU) -> sliced 1x1 expand (+bias, h-sigmoid) -> channel-wise scale;
//   * sliced fully connected layer, DynamicLinear (dynamic_op.py:115-136): y = x W[:out,:in]^T + b[:out];
//   * strided elastic depthwise (DynamicSeparableConv2d with stride 2, dynamic_op.py:73-84), forward, data
//     gradient and filter gradient on an explicit [C][ks*ks] active filter.
#include "ofa_common.cuh"
#include "kernels.h"

namespace ofa {

// the epilogue activations plus the squeeze-and-excite gate; h-sigmoid lives here only, so the hot kernels' epilogue
// switch (apply_act in ofa_common.cuh) stays at three cases
__device__ __forceinline__ float apply_act_gate(float v, int act) {
  if (act == OFA_ACT_HSIGMOID) return fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
  return apply_act(v, act);
}

// ---- sliced linear: Y[n, o] = act(b[o] + sum_i X[n, i] * W[o * ldw + i]) -------------------------------------
// one warp per output element (the layers are [batch, <= 1280] x [<= 1280]: latency-sized, not bandwidth-sized)
__global__ void linear_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                  long long ldw, const float* __restrict__ bias, int N, int in, int out, int act,
                                  float* __restrict__ y, long long ldy) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N * out) return;
  const int n = warp / out, o = warp - n * out;
  const float* xr = x + (size_t)n * ldx;
  const float* wr = w + (size_t)o * ldw;
  float acc = 0.f;
  for (int i = lane; i < in; i += 32) acc = fmaf(xr[i], wr[i], acc);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) y[(size_t)n * ldy + o] = apply_act_gate(acc + (bias ? bias[o] : 0.f), act);
}

// dX[n, i] = sum_o dZ[n, o] * W[o * ldw + i]: thread per (n, i), coalesced over i
__global__ void linear_bwd_data_kernel(const float* __restrict__ dz, long long lddz, const float* __restrict__ w,
                                       long long ldw, int N, int in, int out, float* __restrict__ dx, long long lddx) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * in) return;
  const int n = idx / in, i = idx - n * in;
  float acc = 0.f;
  for (int o = 0; o < out; ++o) acc = fmaf(dz[(size_t)n * lddz + o], w[(size_t)o * ldw + i], acc);
  dx[(size_t)n * lddx + i] = acc;
}

// dW[o * lddw + i] = sum_n dZ[n, o] * X[n, i] (overwrites the active slice), db[o] = sum_n dZ[n, o]
__global__ void linear_bwd_weight_kernel(const float* __restrict__ dz, long long lddz, const float* __restrict__ x,
                                         long long ldx, int N, int in, int out, float* __restrict__ dw,
                                         long long lddw, float* __restrict__ db) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= out * (in + 1)) return;
  const int o = idx / (in + 1), i = idx - o * (in + 1);
  float acc = 0.f;
  if (i < in) {
    for (int n = 0; n < N; ++n) acc = fmaf(dz[(size_t)n * lddz + o], x[(size_t)n * ldx + i], acc);
    dw[(size_t)o * lddw + i] = acc;
  } else if (db) {
    for (int n = 0; n < N; ++n) acc += dz[(size_t)n * lddz + o];
    db[o] = acc;
  }
}

// dZ = dY * act'(z) with the derivative read off the OUTPUT y (ReLU: y > 0; h-sigmoid: 0 < y < 1 -> 1/6)
__global__ void act_bwd_from_output_kernel(const float* __restrict__ dy, const float* __restrict__ y, int act,
                                           float* __restrict__ dz, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float v = y[i];
  float g = 1.f;
  if (act == OFA_ACT_RELU) g = v > 0.f ? 1.f : 0.f;
  else if (act == OFA_ACT_RELU6) g = (v > 0.f && v < 6.f) ? 1.f : 0.f;
  else if (act == OFA_ACT_HSIGMOID) g = (v > 0.f && v < 1.f) ? (1.f / 6.f) : 0.f;
  dz[i] = dy[i] * g;
}

int launch_linear_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* bias, int N, int in,
                      int out, int act, float* y, long long ldy, cudaStream_t st) {
  const long long warps = (long long)N * out;
  if (warps == 0) return OFA_OK;
  linear_fwd_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(x, ldx, w, ldw, bias, N, in, out, act, y, ldy);
  return check_launch("linear_fwd_kernel");
}
int launch_linear_bwd_data(const float* dz, long long lddz, const float* w, long long ldw, int N, int in, int out,
                           float* dx, long long lddx, cudaStream_t st) {
  const long long t = (long long)N * in;
  if (t == 0) return OFA_OK;
  linear_bwd_data_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(dz, lddz, w, ldw, N, in, out, dx, lddx);
  return check_launch("linear_bwd_data_kernel");
}
int launch_linear_bwd_weight(const float* dz, long long lddz, const float* x, long long ldx, int N, int in, int out,
                             float* dw, long long lddw, float* db, cudaStream_t st) {
  const long long t = (long long)out * (in + 1);
  if (t == 0) return OFA_OK;
  linear_bwd_weight_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(dz, lddz, x, ldx, N, in, out, dw, lddw, db);
  return check_launch("linear_bwd_weight_kernel");
}
int launch_act_bwd_from_output(const float* dy, const float* y, int act, float* dz, long long total, cudaStream_t st) {
  if (total == 0) return OFA_OK;
  act_bwd_from_output_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dy, y, act, dz, total);
  return check_launch("act_bwd_from_output_kernel");
}

// ---- squeeze-and-excite pieces --------------------------------------------------------------------------
// pooled[n, c] = mean over (h, w) of x;  with dy: ds[n, c] = sum over (h, w) of x * dy.  One block per (n, c).
__global__ void __launch_bounds__(128)
plane_reduce_kernel(TV x, TV dy, int with_dy, float scale, float* __restrict__ out) {
  __shared__ float red[4];
  const int n = blockIdx.x / x.c, c = blockIdx.x - n * x.c;
  const int HW = x.h * x.w;
  float acc = 0.f;
  for (int p = threadIdx.x; p < HW; p += 128) {
    const int h = p / x.w, w = p - h * x.w;
    const float v = x.ld(x.off(n, c, h, w));
    acc += with_dy ? v * dy.ld(dy.off(n, c, h, w)) : v;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (red[0] + red[1] + red[2] + red[3]) * scale;
}

// y = x * s[n, c]  (+ add[n, c] when given: the pooled branch's gradient, already divided by H*W)
__global__ void channel_scale_kernel(TV x, TV y, const float* __restrict__ s, const float* __restrict__ add,
                                     long long total) {
  const long long HW = (long long)x.h * x.w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // thread index runs over the memory order of x when it is NHWC (c innermost), else NCHW order
    int n, c, h, w;
    if (x.sc == 1) {
      c = (int)(i % x.c);
      long long r = i / x.c;
      w = (int)(r % x.w); r /= x.w;
      h = (int)(r % x.h); n = (int)(r / x.h);
    } else {
      w = (int)(i % x.w);
      long long r = i / x.w;
      h = (int)(r % x.h); r /= x.h;
      c = (int)(r % x.c); n = (int)(r / x.c);
    }
    (void)HW;
    float v = x.ld(x.off(n, c, h, w)) * s[(size_t)n * x.c + c];
    if (add) v += add[(size_t)n * x.c + c];
    y.st(y.off(n, c, h, w), v);
  }
}

int launch_plane_reduce(const TV& x, const TV* dy, float scale, float* out, cudaStream_t st) {
  const long long blocks = (long long)x.n * x.c;
  if (blocks == 0) return OFA_OK;
  if (blocks >= (1ll << 31)) return fail(OFA_ERR_UNSUPPORTED, "plane reduce: too many planes");
  plane_reduce_kernel<<<(unsigned)blocks, 128, 0, st>>>(x, dy ? *dy : x, dy ? 1 : 0, scale, out);
  return check_launch("plane_reduce_kernel");
}
int launch_channel_scale(const TV& x, const TV& y, const float* s, const float* add, cudaStream_t st) {
  const long long total = (long long)x.n * x.c * x.h * x.w;
  if (total == 0) return OFA_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  channel_scale_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, s, add, total);
  return check_launch("channel_scale_kernel");
}

// ---- strided depthwise on an explicit active filter [C][ks*ks] -----------------------------------------
// out[n, c, oh, ow] = sum x[n, c, oh*s + ky - R, ow*s + kx - R] * f[c][ky][kx],  OH = (H - 1) / s + 1 (same padding)
__global__ void dw_strided_fwd_kernel(TV x, TV y, const float* __restrict__ f, int ks, int stride, Epi epi,
                                      long long total) {
  const int R = ks / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % y.w);
    long long r = i / y.w;
    const int oh = (int)(r % y.h); r /= y.h;
    const int c = (int)(r % y.c), n = (int)(r / y.c);
    const float* fc = f + (size_t)c * ks * ks;
    float acc = 0.f;
    for (int ky = 0; ky < ks; ++ky) {
      const int ih = oh * stride + ky - R;
      if (ih < 0 || ih >= x.h) continue;
      for (int kx = 0; kx < ks; ++kx) {
        const int iw = ow * stride + kx - R;
        if (iw < 0 || iw >= x.w) continue;
        acc = fmaf(x.ld(x.off(n, c, ih, iw)), fc[ky * ks + kx], acc);
      }
    }
    float sc, sh;
    epi_scale_shift(epi, c, sc, sh);
    y.st(y.off(n, c, oh, ow), apply_act(fmaf(acc, sc, sh), epi.act));
  }
}

// dx[n, c, ih, iw] = sum over (ky, kx) with (ih + R - ky) % s == 0 of dy[n, c, (ih + R - ky) / s, ...] * f[c][ky][kx]
__global__ void dw_strided_bwd_data_kernel(TV dy, TV dx, const float* __restrict__ f, int ks, int stride,
                                           long long total) {
  const int R = ks / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int iw = (int)(i % dx.w);
    long long r = i / dx.w;
    const int ih = (int)(r % dx.h); r /= dx.h;
    const int c = (int)(r % dx.c), n = (int)(r / dx.c);
    const float* fc = f + (size_t)c * ks * ks;
    float acc = 0.f;
    for (int ky = 0; ky < ks; ++ky) {
      const int th = ih + R - ky;
      if (th < 0 || th % stride) continue;
      const int oh = th / stride;
      if (oh >= dy.h) continue;
      for (int kx = 0; kx < ks; ++kx) {
        const int tw = iw + R - kx;
        if (tw < 0 || tw % stride) continue;
        const int ow = tw / stride;
        if (ow >= dy.w) continue;
        acc = fmaf(dy.ld(dy.off(n, c, oh, ow)), fc[ky * ks + kx], acc);
      }
    }
    dx.st(dx.off(n, c, ih, iw), acc);
  }
}

// df[c][ky][kx] = sum over (n, oh, ow) of dy * x[.., oh*s + ky - R, ow*s + kx - R]: one block per (c, tap)
__global__ void __launch_bounds__(128)
dw_strided_bwd_filter_kernel(TV x, TV dy, int ks, int stride, float* __restrict__ df) {
  __shared__ float red[4];
  const int c = blockIdx.x / (ks * ks), tap = blockIdx.x - c * ks * ks;
  const int ky = tap / ks, kx = tap - ky * ks, R = ks / 2;
  const long long P = (long long)dy.n * dy.h * dy.w;
  float acc = 0.f;
  for (long long p = threadIdx.x; p < P; p += 128) {
    const int ow = (int)(p % dy.w);
    long long r = p / dy.w;
    const int oh = (int)(r % dy.h), n = (int)(r / dy.h);
    const int ih = oh * stride + ky - R, iw = ow * stride + kx - R;
    if (ih < 0 || ih >= x.h || iw < 0 || iw >= x.w) continue;
    acc = fmaf(dy.ld(dy.off(n, c, oh, ow)), x.ld(x.off(n, c, ih, iw)), acc);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) df[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
}

static unsigned ew_blocks(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  return (unsigned)(blocks > cap ? cap : blocks);
}

int launch_dw_strided_fwd(const TV& x, const TV& y, const float* f, int ks, int stride, const Epi& epi,
                          cudaStream_t st) {
  const long long total = (long long)y.n * y.c * y.h * y.w;
  if (total == 0) return OFA_OK;
  dw_strided_fwd_kernel<<<ew_blocks(total), 256, 0, st>>>(x, y, f, ks, stride, epi, total);
  return check_launch("dw_strided_fwd_kernel");
}
int launch_dw_strided_bwd_data(const TV& dy, const TV& dx, const float* f, int ks, int stride, cudaStream_t st) {
  const long long total = (long long)dx.n * dx.c * dx.h * dx.w;
  if (total == 0) return OFA_OK;
  dw_strided_bwd_data_kernel<<<ew_blocks(total), 256, 0, st>>>(dy, dx, f, ks, stride, total);
  return check_launch("dw_strided_bwd_data_kernel");
}
int launch_dw_strided_bwd_filter(const TV& x, const TV& dy, int ks, int stride, float* df, cudaStream_t st) {
  const long long blocks = (long long)x.c * ks * ks;
  if (blocks == 0) return OFA_OK;
  dw_strided_bwd_filter_kernel<<<(unsigned)blocks, 128, 0, st>>>(x, dy, ks, stride, df);
  return check_launch("dw_strided_bwd_filter_kernel");
}

}  // namespace ofa
