// CUDA-core kernels of the elastic-MBConv SR path: any layout (strided views), fp32 or bf16 I/O,
// fp32 math.  This is the exact path (fp32 parity, training forward/backward); the tensor-core /
// TMA kernels of the bf16 inference path live in conv_tc.cu and dw_fast.cu.
#include "ofa_common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace ofa {

// =================================================================================================
// (a1) active depthwise filter
// =================================================================================================
// One block per AF_CH channels: the weights and the two transform matrices are staged in shared memory (coalesced
// loads), then the 7 -> 5 step runs over (channel, output tap) pairs and the -> 3 step over (channel, tap) -- the same
// fmaf order as active_filter_channel (ofa_common.cuh), so the values are bit-identical to the per-channel form the
// other kernels use.  (Round 1 ran one THREAD per channel: 625 dependent FMAs, ~13 us per call.)
//   chunked == 0: out[c][tap]                (ofa_dw_active_filter, the planar depthwise kernel)
//   chunked == 1: out[c / 64][tap'][c % 64]  with tap' = ks*ks - 1 - tap when flip (the data-gradient filter): the
//                 layout dw_fast_kernel fetches with one bulk copy per 64-channel chunk
constexpr int AF_CH = 16;

__global__ void __launch_bounds__(256)
active_filter_kernel(const float* __restrict__ w7, int kmax, const float* __restrict__ m75,
                     const float* __restrict__ m53, int transform_on, int ks, int C, int chunked, int flip,
                     float* __restrict__ out) {
  pdl_wait();
  __shared__ float s_w[AF_CH * 49];
  __shared__ float s_m75[625];
  __shared__ float s_m53[81];
  __shared__ float s_k5[AF_CH * 25];
  const int c0 = blockIdx.x * AF_CH;
  const int nch = min(AF_CH, C - c0);
  const int tid = threadIdx.x;
  const int kk = kmax * kmax, T = ks * ks;
  const bool transform = transform_on && ks < kmax;
  const bool step75 = transform && kmax == 7 && m75 != nullptr;
  for (int i = tid; i < nch * kk; i += blockDim.x) s_w[i] = w7[(size_t)c0 * kk + i];
  if (step75) for (int i = tid; i < 625; i += blockDim.x) s_m75[i] = m75[i];
  if (transform && ks == 3) for (int i = tid; i < 81; i += blockDim.x) s_m53[i] = m53[i];
  __syncthreads();
  if (step75) {
    for (int it = tid; it < nch * 25; it += blockDim.x) {
      const int c = it / 25, j = it - c * 25;
      const float* w = s_w + c * 49;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 25; ++i) acc = fmaf(w[(i / 5 + 1) * 7 + (i % 5 + 1)], s_m75[j * 25 + i], acc);
      s_k5[c * 25 + j] = acc;
    }
    __syncthreads();
  }
  const int R = ks / 2;
  for (int it = tid; it < AF_CH * T; it += blockDim.x) {
    const int c = it % AF_CH, j = it / AF_CH;
    if (c >= nch) continue;
    const float* w = s_w + c * kk;
    float v;
    if (!transform) {
      const int s = kmax / 2 - R;
      v = w[(j / ks + s) * kmax + (j % ks + s)];
    } else {
      const float* cur = step75 ? (s_k5 + c * 25) : w;
      const int kc = step75 ? 5 : kmax;
      if (ks == kc) {
        v = cur[j];
      } else {  // ks == 3
        const int s = kc / 2 - 1;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) acc = fmaf(cur[(i / 3 + s) * kc + (i % 3 + s)], s_m53[j * 9 + i], acc);
        v = acc;
      }
    }
    const int cg = c0 + c;
    if (chunked) out[((size_t)(cg >> 6) * T + (flip ? T - 1 - j : j)) * 64 + (cg & 63)] = v;
    else out[(size_t)cg * T + j] = v;
  }
}

int launch_active_filter(const float* w7, int kmax, const float* m75, const float* m53,
                         int transform_on, int ks, int C, float* out, cudaStream_t st) {
  if (C == 0) return OFA_OK;
  launch_pdl(active_filter_kernel, dim3((C + AF_CH - 1) / AF_CH), dim3(256), 0, st, w7, kmax, m75, m53, transform_on, ks, C, 0, 0,
             out);
  return check_launch("active_filter_kernel");
}

// C % 64 == 0: [C / 64][ks * ks][64], rotated by 180 degrees when flip
int launch_active_filter_chunked(const float* w7, int kmax, const float* m75, const float* m53, int transform_on, int ks,
                                 int C, int flip, float* out, cudaStream_t st) {
  if (C == 0) return OFA_OK;
  launch_pdl(active_filter_kernel, dim3((C + AF_CH - 1) / AF_CH), dim3(256), 0, st, w7, kmax, m75, m53, transform_on, ks, C, 1,
             flip, out);
  return check_launch("active_filter_kernel");
}

// =================================================================================================
// (a2) depthwise conv, generic layout.  Block = 32 channels x DW_PIX pixels per pass; the block's
// 32 active filters are derived once in the prologue (the transform is applied on the fly, no
// global scratch).  flip = 1 gives the data gradient.
// =================================================================================================
constexpr int DW_CH = 32;
constexpr int DW_THREADS = 256;
constexpr int DW_PIX_PER_BLOCK = 1024;

template <int KS>
__global__ void __launch_bounds__(DW_THREADS)
dw_simt_kernel(TV x, TV y, const float* __restrict__ w7, int kmax, const float* __restrict__ m75,
               const float* __restrict__ m53, int transform_on, int flip, Epi epi, int c_is_inner) {
  __shared__ float sf[DW_CH][KS * KS + 1];
  const int c0 = blockIdx.y * DW_CH;
  const int C = x.c;
  // prologue: one thread per channel derives its filter
  if (threadIdx.x < DW_CH) {
    int c = c0 + threadIdx.x;
    if (c < C) {
      float f[49], tmp[25];
      active_filter_channel(w7 + (size_t)c * kmax * kmax, kmax, m75, m53, transform_on, KS, f, tmp);
      for (int j = 0; j < KS * KS; ++j) sf[threadIdx.x][flip ? (KS * KS - 1 - j) : j] = f[j];
    }
  }
  __syncthreads();

  int cl, pl, pstep;
  if (c_is_inner) { cl = threadIdx.x % DW_CH; pl = threadIdx.x / DW_CH; }
  else            { pl = threadIdx.x % (DW_THREADS / DW_CH); cl = threadIdx.x / (DW_THREADS / DW_CH); }
  pstep = DW_THREADS / DW_CH;
  const int c = c0 + cl;
  if (c >= C) return;
  float scale, shift;
  epi_scale_shift(epi, c, scale, shift);
  const long long HW = (long long)x.h * x.w;
  const long long P = HW * x.n;
  const long long p_begin = (long long)blockIdx.x * DW_PIX_PER_BLOCK;
  long long p_end = p_begin + DW_PIX_PER_BLOCK;
  if (p_end > P) p_end = P;
  constexpr int R = KS / 2;
  for (long long p = p_begin + pl; p < p_end; p += pstep) {
    int n = (int)(p / HW);
    int rem = (int)(p - (long long)n * HW);
    int h = rem / x.w, w = rem - h * x.w;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      int ih = h + ky - R;
      if (ih < 0 || ih >= x.h) continue;
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        int iw = w + kx - R;
        if (iw < 0 || iw >= x.w) continue;
        acc = fmaf(x.ld(x.off(n, c, ih, iw)), sf[cl][ky * KS + kx], acc);
      }
    }
    float v = apply_act(fmaf(acc, scale, shift), epi.act);
    if (epi.res.ptr) v += epi.res.ld(epi.res.off(n, c, h, w));
    y.st(y.off(n, c, h, w), v);
  }
}

int launch_dw_simt(const TV& x, const TV& y, const float* w7, int kmax, const float* m75,
                   const float* m53, int transform_on, int ks, int flip, const Epi& epi,
                   cudaStream_t st) {
  long long P = (long long)x.n * x.h * x.w;
  if (P == 0 || x.c == 0) return OFA_OK;
  dim3 grid((unsigned)((P + DW_PIX_PER_BLOCK - 1) / DW_PIX_PER_BLOCK), (x.c + DW_CH - 1) / DW_CH);
  int ci = (x.sc == 1);
  switch (ks) {
    case 3: dw_simt_kernel<3><<<grid, DW_THREADS, 0, st>>>(x, y, w7, kmax, m75, m53, transform_on, flip, epi, ci); break;
    case 5: dw_simt_kernel<5><<<grid, DW_THREADS, 0, st>>>(x, y, w7, kmax, m75, m53, transform_on, flip, epi, ci); break;
    case 7: dw_simt_kernel<7><<<grid, DW_THREADS, 0, st>>>(x, y, w7, kmax, m75, m53, transform_on, flip, epi, ci); break;
    default: return fail(OFA_ERR_UNSUPPORTED, "depthwise kernel size %d (supported: 3, 5, 7)", ks);
  }
  return check_launch("dw_simt_kernel");
}

// =================================================================================================
// (a3, a4, a9-a11) dense conv as an implicit GEMM on CUDA cores.
//   C[p, o] = sum_k A[p, k] * B[k, o],  k = (ky*ks + kx)*cin + ci,  A gathered with zero padding.
//   Tile 64 pixels x 64 outputs, 256 threads, 4x4 register tile, K chunk 16.
// =================================================================================================
constexpr int CV_TP = 64, CV_TO = 64, CV_TK = 16, CV_THREADS = 256;

__global__ void __launch_bounds__(CV_THREADS)
conv_simt_kernel(TV x, TV y, const float* __restrict__ w, long long w_so, long long w_si,
                 long long w_sh, long long w_sw, int cin, int cout, int ks, int flip, int store,
                 Epi epi) {
  __shared__ float sA[CV_TK][CV_TP + 4];
  __shared__ float sB[CV_TK][CV_TO + 4];
  const int H = x.h, W = x.w;
  const long long HW = (long long)H * W;
  const long long P = HW * x.n;
  const long long p0 = (long long)blockIdx.x * CV_TP;
  const int o0 = blockIdx.y * CV_TO;
  const int K = ks * ks * cin;
  const int R = ks / 2;
  const int tid = threadIdx.x;
  const int tp = (tid % 16) * 4;  // pixel sub-tile
  const int to = (tid / 16) * 4;  // output sub-tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // A loader: thread -> (kk = tid % 16, pixel = tid / 16 + 16*r)
  const int a_kk = tid % CV_TK;
  const int a_p = tid / CV_TK;
  int a_n[4], a_h[4], a_w[4];
  bool a_ok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    long long p = p0 + a_p + 16 * r;
    a_ok[r] = p < P;
    long long pp = a_ok[r] ? p : 0;
    a_n[r] = (int)(pp / HW);
    int rem = (int)(pp - (long long)a_n[r] * HW);
    a_h[r] = rem / W;
    a_w[r] = rem - a_h[r] * W;
  }
  // B loader: thread -> (oo = tid % 64, kk = tid / 64 + 4*r)
  const int b_o = tid % CV_TO;
  const int b_k = tid / CV_TO;

  for (int k0 = 0; k0 < K; k0 += CV_TK) {
    {
      int k = k0 + a_kk;
      int tap = k / cin, ci = k - tap * cin;
      int ky = tap / ks, kx = tap - ky * ks;
      bool kok = k < K;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float v = 0.f;
        int ih = a_h[r] + ky - R, iw = a_w[r] + kx - R;
        if (kok && a_ok[r] && ih >= 0 && ih < H && iw >= 0 && iw < W) v = x.ld(x.off(a_n[r], ci, ih, iw));
        sA[a_kk][a_p + 16 * r] = v;
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int kk = b_k + 4 * r;
      int k = k0 + kk;
      float v = 0.f;
      int o = o0 + b_o;
      if (k < K && o < cout) {
        int tap = k / cin, ci = k - tap * cin;
        int ky = tap / ks, kx = tap - ky * ks;
        if (flip) { ky = ks - 1 - ky; kx = ks - 1 - kx; }
        v = w[o * w_so + ci * w_si + ky * w_sh + kx * w_sw];
      }
      sB[kk][b_o] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CV_TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][tp + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][to + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long p = p0 + tp + i;
    if (p >= P) continue;
    int n = (int)(p / HW);
    int rem = (int)(p - (long long)n * HW);
    int h = rem / W, wq = rem - h * W;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int o = o0 + to + j;
      if (o >= cout) continue;
      float scale, shift;
      epi_scale_shift(epi, o, scale, shift);
      float v = apply_act(fmaf(acc[i][j], scale, shift), epi.act);
      int oc, oh, ow;
      store_coord(store, o, h, wq, oc, oh, ow);
      if (epi.res.ptr) v += epi.res.ld(epi.res.off(n, oc, oh, ow));
      y.st(y.off(n, oc, oh, ow), v);
    }
  }
}

int launch_conv_simt(const TV& x, const TV& y, const float* w, long long w_so, long long w_si,
                     long long w_sh, long long w_sw, int cin, int cout, int ks, int flip, int store,
                     const Epi& epi, cudaStream_t st) {
  long long P = (long long)x.n * x.h * x.w;
  if (P == 0 || cout == 0) return OFA_OK;
  dim3 grid((unsigned)((P + CV_TP - 1) / CV_TP), (cout + CV_TO - 1) / CV_TO);
  conv_simt_kernel<<<grid, CV_THREADS, 0, st>>>(x, y, w, w_so, w_si, w_sh, w_sw, cin, cout, ks, flip, store, epi);
  return check_launch("conv_simt_kernel");
}

// =================================================================================================
// weight packing: fp32 strided slice -> bf16 [tap][cout_pad][cin_pad]
// =================================================================================================
__global__ void pack_weight_kernel(const float* __restrict__ w, long long w_so, long long w_si,
                                   long long w_sh, long long w_sw, int cin, int cout, int ks,
                                   int cin_pad, int cout_pad, int store, int f16,
                                   uint16_t* __restrict__ out) {
  pdl_wait();
  long long total = (long long)ks * ks * cout_pad * cin_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ci = (int)(i % cin_pad);
    long long r = i / cin_pad;
    int o = (int)(r % cout_pad);
    int tap = (int)(r / cout_pad);
    int ky = tap / ks, kx = tap - ky * ks;
    float v = 0.f;
    if (ci < cin && o < cout) {
      // PixelShuffle layers are packed sub-pixel major (row s*q + c' holds conv channel 4c' + s) so a
      // tile's accumulator columns of one sub-pixel are contiguous output channels
      int oo = o;
      if (store == OFA_STORE_PIXELSHUFFLE2) { int q = cout >> 2; oo = 4 * (o % q) + o / q; }
      v = w[oo * w_so + ci * w_si + ky * w_sh + kx * w_sw];
    }
    out[i] = cvt16(v, f16);
  }
}

int launch_pack_weight(const float* w, long long w_so, long long w_si, long long w_sh, long long w_sw,
                       int cin, int cout, int ks, int cin_pad, int cout_pad, int store, int f16, void* out,
                       cudaStream_t st) {
  long long total = (long long)ks * ks * cout_pad * cin_pad;
  if (total == 0) return OFA_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  launch_pdl(pack_weight_kernel, dim3(blocks), dim3(256), 0, st, w, w_so, w_si, w_sh, w_sw, cin, cout, ks, cin_pad, cout_pad, store, f16,
                                             reinterpret_cast<uint16_t*>(out));
  return check_launch("pack_weight_kernel");
}

// the same pack for a device-resident list of jobs (blockIdx.y = job): one launch refreshes every 16-bit weight copy a
// forward pass is going to use
__global__ void pack_weights_multi_kernel(const OfaPackJob* __restrict__ jobs) {
  pdl_wait();
  const OfaPackJob j = jobs[blockIdx.y];
  const int ks = j.ks, cin_pad = j.cin_pad, cout_pad = j.cout_pad;
  const int f16 = j.dtype == OFA_F16 ? 1 : 0;
  const float* __restrict__ w = j.w;
  uint16_t* __restrict__ out = reinterpret_cast<uint16_t*>(j.out);
  const int total = ks * ks * cout_pad * cin_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % cin_pad;
    const int r = i / cin_pad;
    const int o = r % cout_pad;
    const int tap = r / cout_pad;
    const int ky = tap / ks, kx = tap - ky * ks;
    float v = 0.f;
    if (ci < j.cin && o < j.cout) {
      int oo = o;
      if (j.store == OFA_STORE_PIXELSHUFFLE2) { const int q = j.cout >> 2; oo = 4 * (o % q) + o / q; }
      v = w[oo * j.w_so + ci * j.w_si + ky * j.w_sh + kx * j.w_sw];
    }
    out[i] = cvt16(v, f16);
  }
}

// up to four jobs passed BY VALUE (kernel parameters): the training block packs its two forward and two
// data-gradient weight copies with one launch and no device-side job table
struct PackJobs4 { OfaPackJob j[4]; };

__device__ __forceinline__ void pack_job(const OfaPackJob& j, int start, int stride) {
  const int ks = j.ks, cin_pad = j.cin_pad, cout_pad = j.cout_pad;
  const int f16 = j.dtype == OFA_F16 ? 1 : 0;
  const float* __restrict__ w = j.w;
  uint16_t* __restrict__ out = reinterpret_cast<uint16_t*>(j.out);
  const int total = ks * ks * cout_pad * cin_pad;
  for (int i = start; i < total; i += stride) {
    const int ci = i % cin_pad;
    const int r = i / cin_pad;
    const int o = r % cout_pad;
    const int tap = r / cout_pad;
    const int ky = tap / ks, kx = tap - ky * ks;
    float v = 0.f;
    if (ci < j.cin && o < j.cout) {
      int oo = o;
      if (j.store == OFA_STORE_PIXELSHUFFLE2) { const int q = j.cout >> 2; oo = 4 * (o % q) + o / q; }
      v = w[oo * j.w_so + ci * j.w_si + ky * j.w_sh + kx * j.w_sw];
    }
    out[i] = cvt16(v, f16);
  }
}

__global__ void pack_weights4_kernel(const __grid_constant__ PackJobs4 jobs) {
  pdl_wait();
  pack_job(jobs.j[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

int launch_pack_weights4(const OfaPackJob* jobs_host, int njobs, cudaStream_t st) {
  if (njobs <= 0) return OFA_OK;
  if (njobs > 4) return fail(OFA_ERR_ARG, "launch_pack_weights4: at most 4 jobs");
  PackJobs4 pj;
  memset(&pj, 0, sizeof(pj));
  for (int i = 0; i < njobs; ++i) pj.j[i] = jobs_host[i];
  launch_pdl(pack_weights4_kernel, dim3(24, (unsigned)njobs), dim3(256), 0, st, pj);
  return check_launch("pack_weights4_kernel");
}

int launch_pack_weights_multi(const OfaPackJob* jobs_device, int njobs, cudaStream_t st) {
  if (njobs <= 0) return OFA_OK;
  launch_pdl(pack_weights_multi_kernel, dim3(dim3(24, (unsigned)njobs)), dim3(256), 0, st, jobs_device);
  return check_launch("pack_weights_multi_kernel");
}

// =================================================================================================
// elementwise: y = store(act(affine(x))) + residual
// =================================================================================================
// -------------------------------------------------------------------------------------------------
// dense-NHWC fast paths: element (pixel p, channel c) lives at p * C + c, so the hot elementwise and
// per-channel reduction kernels index with one multiply (no div/mod chains) and move channel PAIRS.
// -------------------------------------------------------------------------------------------------
inline bool tv_nhwc_dense(const TV& t) {
  return t.sc == 1 && t.sw == t.c && t.sh == (long long)t.w * t.c && t.sn == (long long)t.h * t.w * t.c;
}
__device__ __forceinline__ float2 tv_ld2(const TV& t, long long o) {   // o even, 2 consecutive channels
  if (t.dtype == OFA_F32) return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(t.ptr) + o);
  return unpack16(*reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint16_t*>(t.ptr) + o), t.dtype == OFA_F16);
}
__device__ __forceinline__ void tv_st2(const TV& t, long long o, float2 v) {
  if (t.dtype == OFA_F32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(t.ptr) + o) = v;
  else *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(t.ptr) + o) = pack16(v.x, v.y, t.dtype == OFA_F16);
}
inline bool tv_pair_ok(const TV& t) {
  const int es = t.dtype == OFA_F32 ? 4 : 2;
  return tv_nhwc_dense(t) && (t.c % 2 == 0) && (reinterpret_cast<uintptr_t>(t.ptr) % (2 * es) == 0);
}

__global__ void affine_act_nhwc_kernel(TV x, TV y, Epi epi, unsigned pairs, unsigned cpairs) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += gridDim.x * blockDim.x) {
    const int c = 2 * (int)(i % cpairs);
    float s0, h0, s1, h1;
    epi_scale_shift(epi, c, s0, h0);
    epi_scale_shift(epi, c + 1, s1, h1);
    const long long o = 2ll * i;
    float2 v = tv_ld2(x, o);
    v.x = apply_act(fmaf(v.x, s0, h0), epi.act);
    v.y = apply_act(fmaf(v.y, s1, h1), epi.act);
    if (epi.res.ptr) { const float2 r = tv_ld2(epi.res, o); v.x += r.x; v.y += r.y; }
    tv_st2(y, o, v);
  }
}

// 16-bit dense NHWC, C % 8 == 0: a thread owns one group of 8 channels (one 16-byte vector per pixel) for the whole
// launch -- the grid stride is a multiple of the vectors per pixel -- so the per-channel scale / shift are computed
// once into registers, and a pixel's 8 channels move as one 16-byte load and one 16-byte store.
inline bool tv_vec8_ok(const TV& t) {
  return t.dtype != OFA_F32 && tv_nhwc_dense(t) && (t.c % 8 == 0) && (reinterpret_cast<uintptr_t>(t.ptr) % 16 == 0);
}
inline unsigned vec8_blocks(unsigned long long nvec, unsigned V, int threads, int blocks_per_sm) {
  unsigned a = V, b = (unsigned)threads;
  while (b) { const unsigned t = a % b; a = b; b = t; }
  const unsigned m = V / a;                               // blocks must come in multiples of V / gcd(V, threads)
  unsigned long long blocks = (nvec + threads - 1) / threads;
  const unsigned long long cap = (unsigned long long)sm_count() * blocks_per_sm;   // one resident wave
  if (blocks > cap) blocks = cap / m * m >= m ? cap / m * m : m;
  return (unsigned)((blocks + m - 1) / m * m);
}

template <bool HAS_RES>
__global__ void __launch_bounds__(256)
affine_act_vec8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const uint4* __restrict__ res, Epi epi,
                       int f16, unsigned nvec, unsigned V) {
  pdl_wait();
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = 8 * (int)(gid % V);
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) epi_scale_shift(epi, c0 + k, sc[k], sh[k]);
  const int act = epi.act;
  for (unsigned i = gid; i < nvec; i += gridDim.x * blockDim.x) {
    const uint4 v = x[i];
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (HAS_RES) r = res[i];
    const uint32_t in[4] = {v.x, v.y, v.z, v.w}, rr[4] = {r.x, r.y, r.z, r.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = unpack16(in[k], f16);
      f.x = apply_act(fmaf(f.x, sc[2 * k], sh[2 * k]), act);
      f.y = apply_act(fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]), act);
      if (HAS_RES) { const float2 q = unpack16(rr[k], f16); f.x += q.x; f.y += q.y; }
      o[k] = pack16(f.x, f.y, f16);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// dense-NHWC PixelShuffle(2) / PixelUnshuffle(2) reorder (no affine, no residual): 32-bit index arithmetic, the
// thread index runs over the OUTPUT so stores are coalesced
__global__ void reorder_nhwc_kernel(TV x, TV y, int store, unsigned total) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned oc = i % (unsigned)y.c;
    unsigned r = i / (unsigned)y.c;
    const unsigned ow = r % (unsigned)y.w; r /= (unsigned)y.w;
    const unsigned oh = r % (unsigned)y.h, n = r / (unsigned)y.h;
    unsigned c, h, w;
    if (store == OFA_STORE_PIXELSHUFFLE2) {          // out[n, c', 2h+i, 2w+j] = in[n, 4c'+2i+j, h, w]
      c = 4 * oc + 2 * (oh & 1) + (ow & 1); h = oh >> 1; w = ow >> 1;
    } else {                                          // out[n, 4c+2y+x, h', w'] = in[n, c, 2h'+y, 2w'+x]
      c = oc >> 2; h = 2 * oh + ((oc >> 1) & 1); w = 2 * ow + (oc & 1);
    }
    const long long xo = ((long long)(n * x.h + h) * x.w + w) * x.c + c;
    y.st(i, x.ld(xo));
  }
}

// 16-bit dense NHWC PixelShuffle(2) as 16-byte loads and coalesced 4-byte stores: a thread takes the 8 input channels
// 4c'+s of two neighbouring output channels c' = 2l, 2l+1 (one 16-byte vector) and writes one channel pair to each of
// the four sub-pixels s = 2i+j; a warp's 32 words per sub-pixel are contiguous.  PixelUnshuffle(2) is the inverse:
// four 4-byte loads (one per sub-pixel), one 16-byte store.
__global__ void pixel_shuffle2_vec_kernel(const uint4* __restrict__ x, uint32_t* __restrict__ y, int H, int W, int q2,
                                          unsigned total) {   // q2 = output channel pairs per pixel = C_in / 8
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned l = i % (unsigned)q2;
    unsigned r = i / (unsigned)q2;                           // input pixel index (n, h, w)
    const unsigned w = r % (unsigned)W; r /= (unsigned)W;
    const unsigned h = r % (unsigned)H, n = r / (unsigned)H;
    const uint4 v = x[i];
    const size_t orow = ((size_t)n * 2 * H + 2 * h) * 2 * W + 2 * w;   // output pixel (2h, 2w)
    uint32_t* o = y + orow * q2 + l;
    o[0] = __byte_perm(v.x, v.z, 0x5410);                                 // s = 0: (c'0 s0, c'1 s0)
    o[q2] = __byte_perm(v.x, v.z, 0x7632);                                // s = 1 -> pixel (2h, 2w + 1)
    o[(size_t)2 * W * q2] = __byte_perm(v.y, v.w, 0x5410);                // s = 2 -> pixel (2h + 1, 2w)
    o[(size_t)2 * W * q2 + q2] = __byte_perm(v.y, v.w, 0x7632);           // s = 3
  }
}

__global__ void pixel_unshuffle2_vec_kernel(const uint32_t* __restrict__ x, uint4* __restrict__ y, int Ho, int Wo,
                                            int c2, unsigned total) {   // c2 = input channel pairs per pixel
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned l = i % (unsigned)c2;
    unsigned r = i / (unsigned)c2;                           // output pixel index (n, h', w')
    const unsigned w = r % (unsigned)Wo; r /= (unsigned)Wo;
    const unsigned h = r % (unsigned)Ho, n = r / (unsigned)Ho;
    const size_t irow = ((size_t)n * 2 * Ho + 2 * h) * 2 * Wo + 2 * w;
    const uint32_t* p = x + irow * c2 + l;
    const uint32_t s0 = p[0], s1 = p[c2], s2 = p[(size_t)2 * Wo * c2], s3 = p[(size_t)2 * Wo * c2 + c2];
    y[i] = make_uint4(__byte_perm(s0, s1, 0x5410), __byte_perm(s2, s3, 0x5410), __byte_perm(s0, s1, 0x7632),
                      __byte_perm(s2, s3, 0x7632));
  }
}

__global__ void affine_act_kernel(TV x, TV y, Epi epi, int store, int c_is_inner) {
  const long long total = (long long)x.n * x.c * x.h * x.w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int n, c, h, w;
    long long r = i;
    if (c_is_inner) {
      c = (int)(r % x.c); r /= x.c;
      w = (int)(r % x.w); r /= x.w;
      h = (int)(r % x.h); n = (int)(r / x.h);
    } else {
      w = (int)(r % x.w); r /= x.w;
      h = (int)(r % x.h); r /= x.h;
      c = (int)(r % x.c); n = (int)(r / x.c);
    }
    float scale, shift;
    epi_scale_shift(epi, c, scale, shift);
    float v = apply_act(fmaf(x.ld(x.off(n, c, h, w)), scale, shift), epi.act);
    int oc, oh, ow;
    store_coord(store, c, h, w, oc, oh, ow);
    if (epi.res.ptr) v += epi.res.ld(epi.res.off(n, oc, oh, ow));
    y.st(y.off(n, oc, oh, ow), v);
  }
}

int launch_affine_act(const TV& x, const TV& y, const Epi& epi, int store, cudaStream_t st) {
  long long total = (long long)x.n * x.c * x.h * x.w;
  if (total == 0) return OFA_OK;
  long long blocks = (total + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (store == OFA_STORE_PLAIN && total < (1ll << 34) && tv_vec8_ok(x) && tv_vec8_ok(y) && y.dtype == x.dtype &&
      (!epi.res.ptr || (tv_vec8_ok(epi.res) && epi.res.dtype == x.dtype))) {
    const unsigned nvec = (unsigned)(total / 8), V = (unsigned)(x.c / 8);
    const unsigned nb = vec8_blocks(nvec, V, 256, 8);
    const uint4* xp = reinterpret_cast<const uint4*>(x.ptr);
    uint4* yp = reinterpret_cast<uint4*>(y.ptr);
    if (epi.res.ptr)
      launch_pdl(affine_act_vec8_kernel<true>, dim3(nb), dim3(256), 0, st, xp, yp, reinterpret_cast<const uint4*>(epi.res.ptr), epi,
                                                       x.dtype == OFA_F16, nvec, V);
    else
      launch_pdl(affine_act_vec8_kernel<false>, dim3(nb), dim3(256), 0, st, xp, yp, nullptr, epi, x.dtype == OFA_F16, nvec, V);
    return check_launch("affine_act_vec8_kernel");
  }
  if (store == OFA_STORE_PLAIN && total < (1ll << 32) && tv_pair_ok(x) && tv_pair_ok(y) &&
      (!epi.res.ptr || tv_pair_ok(epi.res))) {
    affine_act_nhwc_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, epi, (unsigned)(total / 2), (unsigned)(x.c / 2));
    return check_launch("affine_act_nhwc_kernel");
  }
  const bool pure_reorder = store != OFA_STORE_PLAIN && total < (1ll << 31) && tv_nhwc_dense(x) && tv_nhwc_dense(y) &&
                            !epi.res.ptr && !epi.gamma && !epi.beta && !epi.mean && !epi.var && epi.act == OFA_ACT_NONE;
  if (pure_reorder && x.dtype != OFA_F32 && y.dtype == x.dtype && x.c % 8 == 0 && store == OFA_STORE_PIXELSHUFFLE2 &&
      reinterpret_cast<uintptr_t>(x.ptr) % 16 == 0 && reinterpret_cast<uintptr_t>(y.ptr) % 4 == 0) {
    const unsigned nv = (unsigned)(total / 8);
    long long nb = (nv + 255) / 256;
    if (nb > (long long)sm_count() * 8) nb = (long long)sm_count() * 8;
    pixel_shuffle2_vec_kernel<<<(unsigned)nb, 256, 0, st>>>(reinterpret_cast<const uint4*>(x.ptr),
                                                            reinterpret_cast<uint32_t*>(y.ptr), x.h, x.w, x.c / 8, nv);
    return check_launch("pixel_shuffle2_vec_kernel");
  }
  if (pure_reorder && x.dtype != OFA_F32 && y.dtype == x.dtype && x.c % 2 == 0 && store == OFA_STORE_PIXELUNSHUFFLE2 &&
      reinterpret_cast<uintptr_t>(y.ptr) % 16 == 0 && reinterpret_cast<uintptr_t>(x.ptr) % 4 == 0) {
    const unsigned nv = (unsigned)(total / 8);
    long long nb = (nv + 255) / 256;
    if (nb > (long long)sm_count() * 8) nb = (long long)sm_count() * 8;
    pixel_unshuffle2_vec_kernel<<<(unsigned)nb, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(x.ptr),
                                                              reinterpret_cast<uint4*>(y.ptr), y.h, y.w, x.c / 2, nv);
    return check_launch("pixel_unshuffle2_vec_kernel");
  }
  if (pure_reorder) {
    reorder_nhwc_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, store, (unsigned)total);
    return check_launch("reorder_nhwc_kernel");
  }
  affine_act_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, epi, store, x.sc == 1);
  return check_launch("affine_act_kernel");
}

// =================================================================================================
// (a5) training BN: batch statistics.  One block per 32 channels, 8 pixel lanes; two passes
// (mean, then centred second moment) so the variance matches F.batch_norm to fp32 rounding.
// =================================================================================================
constexpr int BN_CH = 32, BN_THREADS = 256, BN_PL = BN_THREADS / BN_CH;

__device__ __forceinline__ void pix_decode(long long p, long long HW, int W, int& n, int& h, int& w) {
  n = (int)(p / HW);
  int rem = (int)(p - (long long)n * HW);
  h = rem / W;
  w = rem - h * W;
}

// Batch statistics as a two-level reduction: grid = (channel groups, pixel splits); a block folds its pixel
// range with per-thread Welford updates and Chan's pairwise combination (no catastrophic cancellation when
// mean^2 >> var), writes one (count, mean, M2) triple per channel; a second tiny kernel combines the splits in a
// fixed order in double precision -> deterministic, and every SM takes part.
struct Wf { float n, mean, m2; };
__device__ __forceinline__ void wf_add(Wf& a, float x) {
  a.n += 1.f;
  const float d = x - a.mean;
  a.mean += d / a.n;
  a.m2 = fmaf(d, x - a.mean, a.m2);
}
__device__ __forceinline__ void wf_merge(Wf& a, const Wf& b) {
  if (b.n == 0.f) return;
  const float n = a.n + b.n, d = b.mean - a.mean;
  a.mean += d * (b.n / n);
  a.m2 += b.m2 + d * d * (a.n * b.n / n);
  a.n = n;
}

__global__ void __launch_bounds__(BN_THREADS)
bn_stats_partial_kernel(TV x, float* __restrict__ part, int splits, long long per_split, int c_is_inner) {
  __shared__ Wf red[BN_PL][BN_CH + 1];
  int cl, pl;
  if (c_is_inner) { cl = threadIdx.x % BN_CH; pl = threadIdx.x / BN_CH; }
  else            { pl = threadIdx.x % BN_PL; cl = threadIdx.x / BN_PL; }
  const int c = blockIdx.x * BN_CH + cl;
  const long long HW = (long long)x.h * x.w;
  const long long P = HW * x.n;
  const long long p_lo = (long long)blockIdx.y * per_split;
  const long long p_hi = p_lo + per_split < P ? p_lo + per_split : P;
  Wf w = {0.f, 0.f, 0.f};
  if (c < x.c)
    for (long long p = p_lo + pl; p < p_hi; p += BN_PL) {
      int n, h, ww;
      pix_decode(p, HW, x.w, n, h, ww);
      wf_add(w, x.ld(x.off(n, c, h, ww)));
    }
  red[pl][cl] = w;
  __syncthreads();
  if (threadIdx.x < BN_CH) {
    const int cc = blockIdx.x * BN_CH + threadIdx.x;
    if (cc < x.c) {
      Wf t = red[0][threadIdx.x];
      for (int i = 1; i < BN_PL; ++i) wf_merge(t, red[i][threadIdx.x]);
      float* o = part + ((size_t)cc * gridDim.y + blockIdx.y) * 3;
      o[0] = t.n; o[1] = t.mean; o[2] = t.m2;
    }
  }
}

// dense NHWC: lane = channel PAIR (a block covers 64 channels), 8 pixel lanes, one multiply per element
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_partial_nhwc_kernel(TV x, float* __restrict__ part, long long per_split) {
  __shared__ Wf red[BN_PL][2 * BN_CH + 1];
  const int cl = threadIdx.x % BN_CH, pl = threadIdx.x / BN_CH;
  const int c = (blockIdx.x * BN_CH + cl) * 2;
  const long long P = (long long)x.n * x.h * x.w;
  const long long p_lo = (long long)blockIdx.y * per_split;
  const long long p_hi = p_lo + per_split < P ? p_lo + per_split : P;
  Wf w0 = {0.f, 0.f, 0.f}, w1 = {0.f, 0.f, 0.f};
  if (c < x.c && p_lo + pl < p_hi) {
    // sums of (v - pivot) and (v - pivot)^2 around the thread's first element: no division per element, no
    // cancellation problem (the pivot is a sample of the same channel); converted to (n, mean, M2) once
    const float2 piv = tv_ld2(x, (p_lo + pl) * x.c + c);
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f, cnt = 0.f;
#pragma unroll 4
    for (long long p = p_lo + pl; p < p_hi; p += BN_PL) {
      const float2 v = tv_ld2(x, p * x.c + c);
      const float d0 = v.x - piv.x, d1 = v.y - piv.y;
      s0 += d0; q0 = fmaf(d0, d0, q0);
      s1 += d1; q1 = fmaf(d1, d1, q1);
      cnt += 1.f;
    }
    w0.n = w1.n = cnt;
    w0.mean = piv.x + s0 / cnt; w0.m2 = q0 - s0 * s0 / cnt;
    w1.mean = piv.y + s1 / cnt; w1.m2 = q1 - s1 * s1 / cnt;
  }
  red[pl][2 * cl] = w0;
  red[pl][2 * cl + 1] = w1;
  __syncthreads();
  if (threadIdx.x < 2 * BN_CH) {
    const int cc = blockIdx.x * 2 * BN_CH + threadIdx.x;
    if (cc < x.c) {
      Wf t = red[0][threadIdx.x];
      for (int i = 1; i < BN_PL; ++i) wf_merge(t, red[i][threadIdx.x]);
      float* o = part + ((size_t)cc * gridDim.y + blockIdx.y) * 3;
      o[0] = t.n; o[1] = t.mean; o[2] = t.m2;
    }
  }
}

// 16-bit dense NHWC, C % 8 == 0, C <= 2048: thread = (group of 8 channels, pixel lane), one 16-byte load per pixel;
// blockDim = V * (256 / V) with V = C / 8, so a block spans all channels and 256 / V pixel lanes
__global__ void __launch_bounds__(256)
bn_stats_partial_vec8_kernel(const uint4* __restrict__ x, int f16, int V, long long P, float* __restrict__ part,
                             long long per_split) {
  pdl_wait();
  __shared__ float red[3][2048];
  const int cg = threadIdx.x % V, pl = threadIdx.x / V, PL = blockDim.x / V, C = 8 * V;
  const long long p_lo = (long long)blockIdx.y * per_split;
  const long long p_hi = p_lo + per_split < P ? p_lo + per_split : P;
  // sums of (v - pivot) and (v - pivot)^2 around the block's first pixel (a sample of the same channel: no
  // cancellation problem, no division per element); the pixel lanes then combine by plain sums
  float piv[8], sd[8], sq[8];
  {
    const uint4 v = x[p_lo * V + cg];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = unpack16(w[k], f16); piv[2 * k] = f.x; piv[2 * k + 1] = f.y; }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sd[k] = sq[k] = 0.f;
#pragma unroll 4
  for (long long p = p_lo + pl; p < p_hi; p += PL) {
    const uint4 v = x[p * V + cg];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack16(w[k], f16);
      const float d0 = f.x - piv[2 * k], d1 = f.y - piv[2 * k + 1];
      sd[2 * k] += d0; sq[2 * k] = fmaf(d0, d0, sq[2 * k]);
      sd[2 * k + 1] += d1; sq[2 * k + 1] = fmaf(d1, d1, sq[2 * k + 1]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { red[0][pl * C + 8 * cg + k] = sd[k]; red[1][pl * C + 8 * cg + k] = sq[k]; }
  if (pl == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[2][8 * cg + k] = piv[k];
  }
  __syncthreads();
  const float cnt = (float)(p_hi - p_lo);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float S = 0.f, Q = 0.f;
    for (int i = 0; i < PL; ++i) { S += red[0][i * C + c]; Q += red[1][i * C + c]; }
    float* o = part + ((size_t)c * gridDim.y + blockIdx.y) * 3;
    o[0] = cnt; o[1] = red[2][c] + S / cnt; o[2] = Q - S * S / cnt;
  }
}

// thin tensors (C <= 8: the 3-channel images): every thread walks pixels and keeps all channels in registers —
// the 32-channels-per-block mapping above would leave 29 of 32 lanes idle
constexpr int BN_SMALLC = 8;

__global__ void __launch_bounds__(BN_THREADS)
bn_stats_partial_smallc_kernel(TV x, float* __restrict__ part, long long per_split) {
  __shared__ Wf red[BN_THREADS / 32][BN_SMALLC];
  const long long HW = (long long)x.h * x.w;
  const long long P = HW * x.n;
  const long long p_lo = (long long)blockIdx.y * per_split;
  const long long p_hi = p_lo + per_split < P ? p_lo + per_split : P;
  Wf w[BN_SMALLC];
#pragma unroll
  for (int c = 0; c < BN_SMALLC; ++c) w[c] = {0.f, 0.f, 0.f};
  for (long long p = p_lo + threadIdx.x; p < p_hi; p += BN_THREADS) {
    int n, h, ww;
    pix_decode(p, HW, x.w, n, h, ww);
#pragma unroll
    for (int c = 0; c < BN_SMALLC; ++c)
      if (c < x.c) wf_add(w[c], x.ld(x.off(n, c, h, ww)));
  }
#pragma unroll
  for (int c = 0; c < BN_SMALLC; ++c) {
    for (int o = 16; o > 0; o >>= 1) {
      Wf t;
      t.n = __shfl_down_sync(0xffffffffu, w[c].n, o);
      t.mean = __shfl_down_sync(0xffffffffu, w[c].mean, o);
      t.m2 = __shfl_down_sync(0xffffffffu, w[c].m2, o);
      wf_merge(w[c], t);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = w[c];
  }
  __syncthreads();
  if ((int)threadIdx.x < x.c) {
    Wf t = red[0][threadIdx.x];
    for (int i = 1; i < BN_THREADS / 32; ++i) wf_merge(t, red[i][threadIdx.x]);
    float* o = part + ((size_t)threadIdx.x * gridDim.y + blockIdx.y) * 3;
    o[0] = t.n; o[1] = t.mean; o[2] = t.m2;
  }
}

// one warp per channel: lanes stride over the splits, fixed-order shuffle tree -> deterministic; sums in double,
// no serial chain of divisions (the splits can number a few hundred)
constexpr int BN_FIN_THREADS = 256;
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(BN_FIN_THREADS)
bn_stats_final_kernel(const float* __restrict__ part, int splits, int C, float* __restrict__ mean,
                      float* __restrict__ var, float* __restrict__ rm, float* __restrict__ rv, float momentum, float unbias,
                      long long* __restrict__ num_batches_tracked) {
  pdl_wait();
  const int c = blockIdx.x * (BN_FIN_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  // nn.BatchNorm2d's counter, bumped in the same launch (`bn.num_batches_tracked += 1`, dynamic_op.py:156)
  if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;
  if (c >= C) return;
  double n = 0.0, sm = 0.0;
  for (int s = lane; s < splits; s += 32) {
    const float* o = part + ((size_t)c * splits + s) * 3;
    n += (double)o[0];
    sm += (double)o[0] * (double)o[1];
  }
  n = warp_sum_d(n);
  sm = warp_sum_d(sm);
  const double m = n > 0.0 ? sm / n : 0.0;
  double m2 = 0.0;
  for (int s = lane; s < splits; s += 32) {
    const float* o = part + ((size_t)c * splits + s) * 3;
    const double d = (double)o[1] - m;
    m2 += (double)o[2] + (double)o[0] * d * d;
  }
  m2 = warp_sum_d(m2);
  if (lane == 0) {
    const float mf = (float)m, vf = (float)(n > 0.0 ? m2 / n : 0.0);
    mean[c] = mf;
    var[c] = vf;
    if (rm) {   // running = (1 - momentum) * running + momentum * {mean, unbiased var} on the active slice, same launch
      rm[c] = (1.f - momentum) * rm[c] + momentum * mf;
      rv[c] = (1.f - momentum) * rv[c] + momentum * (vf * unbias);
    }
  }
}

// Stream-ordered scratch (cudaMallocAsync) for the split reductions.  The default memory pool releases its memory
// back to the OS at every synchronisation (release threshold 0); re-creating ~100 small allocations per training
// step then costs hundreds of milliseconds.  Once per device the threshold is raised so the pool keeps its pages.
void keep_async_pool_resident() {
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done[dev] = true;
}

static bool bn_vec8_ok(const TV& t) {
  return tv_vec8_ok(t) && t.c <= 2048 && (long long)t.n * t.h * t.w * (t.c / 8) < (1ll << 40);
}

// pixel splits of a per-channel reduction: 8 blocks of 256 threads per SM (the reductions are latency-bound loads,
// so they want every warp slot), at least 64 pixels each
static int bn_splits(const TV& x, long long* per_split, int vec8_blocks_per_sm) {
  const long long P = (long long)x.n * x.h * x.w;
  const bool v8 = bn_vec8_ok(x);
  const int ch_per_block = (x.c <= BN_SMALLC || v8) ? x.c : tv_pair_ok(x) ? 2 * BN_CH : BN_CH;
  const int groups = (x.c + ch_per_block - 1) / ch_per_block;
  // one resident wave of blocks (a second, partial wave costs as much as the first on these short kernels)
  long long want = ((long long)(v8 ? vec8_blocks_per_sm : 8) * sm_count() + groups - 1) / groups;
  long long maxs = (P + 63) / 64;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  *per_split = (P + want - 1) / want;
  return (int)((P + *per_split - 1) / *per_split);
}

int launch_bn_stats(const TV& x, float* mean, float* var, cudaStream_t st) {
  return launch_bn_stats_update(x, mean, var, nullptr, nullptr, 0.f, nullptr, st);
}

// batch statistics + (rm != nullptr) the running-statistics update and the num_batches_tracked bump in the finalize launch
int launch_bn_stats_update(const TV& x, float* mean, float* var, float* rm, float* rv, float momentum,
                           long long* num_batches_tracked, cudaStream_t st) {
  if (x.c == 0) return OFA_OK;
  const long long count = (long long)x.n * x.h * x.w;
  const float unbias = count > 1 ? (float)((double)count / (double)(count - 1)) : 1.f;
  long long per_split = 0;
  const int splits = bn_splits(x, &per_split, 4);
  float* part = nullptr;
  keep_async_pool_resident();
  OFA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&part), (size_t)splits * x.c * 3 * sizeof(float), st));
  int rc;
  if (x.c <= BN_SMALLC) {
    dim3 grid(1, splits);
    bn_stats_partial_smallc_kernel<<<grid, BN_THREADS, 0, st>>>(x, part, per_split);
    rc = check_launch("bn_stats_partial_smallc_kernel");
  } else if (bn_vec8_ok(x)) {
    const int V = x.c / 8;
    dim3 grid(1, splits);
    launch_pdl(bn_stats_partial_vec8_kernel, dim3(grid), dim3(V * (256 / V)), 0, st, reinterpret_cast<const uint4*>(x.ptr),
                                                                x.dtype == OFA_F16, V, (long long)x.n * x.h * x.w,
                                                                part, per_split);
    rc = check_launch("bn_stats_partial_vec8_kernel");
  } else if (tv_pair_ok(x)) {
    dim3 grid((x.c + 2 * BN_CH - 1) / (2 * BN_CH), splits);
    bn_stats_partial_nhwc_kernel<<<grid, BN_THREADS, 0, st>>>(x, part, per_split);
    rc = check_launch("bn_stats_partial_nhwc_kernel");
  } else {
    dim3 grid((x.c + BN_CH - 1) / BN_CH, splits);
    bn_stats_partial_kernel<<<grid, BN_THREADS, 0, st>>>(x, part, splits, per_split, x.sc == 1);
    rc = check_launch("bn_stats_partial_kernel");
  }
  if (!rc) {
    launch_pdl(bn_stats_final_kernel, dim3((x.c + BN_FIN_THREADS / 32 - 1) / (BN_FIN_THREADS / 32)), dim3(BN_FIN_THREADS), 0, st, 
        part, splits, x.c, mean, var, rm, rv, momentum, unbias, num_batches_tracked);
    rc = check_launch("bn_stats_final_kernel");
  }
  cudaFreeAsync(part, st);
  return rc;
}

__global__ void bn_update_running_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                         float unbias, float* __restrict__ rm, float* __restrict__ rv,
                                         float momentum, int C, long long* __restrict__ num_batches_tracked) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;   // nn.BatchNorm2d's counter, bumped in the same launch
  if (c >= C) return;
  rm[c] = (1.f - momentum) * rm[c] + momentum * mean[c];
  rv[c] = (1.f - momentum) * rv[c] + momentum * (var[c] * unbias);
}

int launch_bn_update_running(const float* mean, const float* var, long long count, float* rm, float* rv,
                             float momentum, int C, long long* num_batches_tracked, cudaStream_t st) {
  if (C == 0) return OFA_OK;
  float unbias = count > 1 ? (float)((double)count / (double)(count - 1)) : 1.f;
  bn_update_running_kernel<<<(C + 127) / 128, 128, 0, st>>>(mean, var, unbias, rm, rv, momentum, C, num_batches_tracked);
  return check_launch("bn_update_running_kernel");
}

// =================================================================================================
// (a14) BN (+act) backward
// =================================================================================================
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_partial_kernel(TV x, TV dy, const float* __restrict__ gamma, const float* __restrict__ beta,
                             const float* __restrict__ mean, const float* __restrict__ var, float eps, int act,
                             float* __restrict__ part, long long per_split, int c_is_inner) {
  __shared__ float red0[BN_PL][BN_CH + 1];
  __shared__ float red1[BN_PL][BN_CH + 1];
  int cl, pl;
  if (c_is_inner) { cl = threadIdx.x % BN_CH; pl = threadIdx.x / BN_CH; }
  else            { pl = threadIdx.x % BN_PL; cl = threadIdx.x / BN_PL; }
  const int c = blockIdx.x * BN_CH + cl;
  const long long HW = (long long)x.h * x.w;
  const long long P = HW * x.n;
  const long long p_lo = (long long)blockIdx.y * per_split;
  const long long p_hi = p_lo + per_split < P ? p_lo + per_split : P;
  float s0 = 0.f, s1 = 0.f;
  if (c < x.c) {
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    const float m = mean[c], rstd = rsqrtf(var[c] + eps);
    for (long long p = p_lo + pl; p < p_hi; p += BN_PL) {
      int n, h, w;
      pix_decode(p, HW, x.w, n, h, w);
      float xhat = (x.ld(x.off(n, c, h, w)) - m) * rstd;
      float z = fmaf(g, xhat, b);
      float dz = dy.ld(dy.off(n, c, h, w)) * act_grad(z, act);
      s0 += dz;
      s1 = fmaf(dz, xhat, s1);
    }
  }
  red0[pl][cl] = s0;
  red1[pl][cl] = s1;
  __syncthreads();
  if (threadIdx.x < BN_CH) {
    int cc = blockIdx.x * BN_CH + threadIdx.x;
    if (cc < x.c) {
      double t0 = 0.0, t1 = 0.0;
      for (int i = 0; i < BN_PL; ++i) { t0 += red0[i][threadIdx.x]; t1 += red1[i][threadIdx.x]; }
      float* o = part + ((size_t)cc * gridDim.y + blockIdx.y) * 2;
      o[0] = (float)t0;
      o[1] = (float)t1;
    }
  }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_partial_nhwc_kernel(TV x, TV dy, const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ mean, const float* __restrict__ var, float eps, int act,
                                  float* __restrict__ part, long long per_split) {
  __shared__ float red0[BN_PL][2 * BN_CH + 1];
  __shared__ float red1[BN_PL][2 * BN_CH + 1];
  const int cl = threadIdx.x % BN_CH, pl = threadIdx.x / BN_CH;
  const int c = (blockIdx.x * BN_CH + cl) * 2;
  const long long P = (long long)x.n * x.h * x.w;
  const long long p_lo = (long long)blockIdx.y * per_split;
  const long long p_hi = p_lo + per_split < P ? p_lo + per_split : P;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  if (c < x.c) {
    const float g0 = gamma ? gamma[c] : 1.f, g1 = gamma ? gamma[c + 1] : 1.f;
    const float be0 = beta ? beta[c] : 0.f, be1 = beta ? beta[c + 1] : 0.f;
    const float m0 = mean[c], m1 = mean[c + 1];
    const float r0 = rsqrtf(var[c] + eps), r1 = rsqrtf(var[c + 1] + eps);
#pragma unroll 4
    for (long long p = p_lo + pl; p < p_hi; p += BN_PL) {
      const long long o = p * x.c + c;
      const float2 xv = tv_ld2(x, o), gv = tv_ld2(dy, o);
      const float xh0 = (xv.x - m0) * r0, xh1 = (xv.y - m1) * r1;
      const float dz0 = gv.x * act_grad(fmaf(g0, xh0, be0), act), dz1 = gv.y * act_grad(fmaf(g1, xh1, be1), act);
      a0 += dz0; a1 += dz1;
      b0 = fmaf(dz0, xh0, b0); b1 = fmaf(dz1, xh1, b1);
    }
  }
  red0[pl][2 * cl] = a0; red0[pl][2 * cl + 1] = a1;
  red1[pl][2 * cl] = b0; red1[pl][2 * cl + 1] = b1;
  __syncthreads();
  if (threadIdx.x < 2 * BN_CH) {
    const int cc = blockIdx.x * 2 * BN_CH + threadIdx.x;
    if (cc < x.c) {
      double t0 = 0.0, t1 = 0.0;
      for (int i = 0; i < BN_PL; ++i) { t0 += red0[i][threadIdx.x]; t1 += red1[i][threadIdx.x]; }
      float* o = part + ((size_t)cc * gridDim.y + blockIdx.y) * 2;
      o[0] = (float)t0;
      o[1] = (float)t1;
    }
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_partial_vec8_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, int f16, int V,
                                  long long P, const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ mean, const float* __restrict__ var, float eps, int act,
                                  float* __restrict__ part, long long per_split) {
  pdl_wait();
  __shared__ float red[2][2048];
  const int cg = threadIdx.x % V, pl = threadIdx.x / V, PL = blockDim.x / V, C = 8 * V;
  const long long p_lo = (long long)blockIdx.y * per_split;
  const long long p_hi = p_lo + per_split < P ? p_lo + per_split : P;
  float rstd[8], nm[8], g[8], b[8], s0[8], s1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = 8 * cg + k;
    g[k] = gamma ? gamma[c] : 1.f;
    b[k] = beta ? beta[c] : 0.f;
    rstd[k] = rsqrtf(var[c] + eps);
    nm[k] = -mean[c] * rstd[k];
    s0[k] = s1[k] = 0.f;
  }
#pragma unroll 2
  for (long long p = p_lo + pl; p < p_hi; p += PL) {
    const uint4 xv = x[p * V + cg], gv = dy[p * V + cg];
    const uint32_t xi[4] = {xv.x, xv.y, xv.z, xv.w}, gi[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 xf = unpack16(xi[k], f16), gf = unpack16(gi[k], f16);
      const float h0 = fmaf(xf.x, rstd[2 * k], nm[2 * k]), h1 = fmaf(xf.y, rstd[2 * k + 1], nm[2 * k + 1]);
      const float d0 = gf.x * act_grad(fmaf(g[2 * k], h0, b[2 * k]), act);
      const float d1 = gf.y * act_grad(fmaf(g[2 * k + 1], h1, b[2 * k + 1]), act);
      s0[2 * k] += d0; s1[2 * k] = fmaf(d0, h0, s1[2 * k]);
      s0[2 * k + 1] += d1; s1[2 * k + 1] = fmaf(d1, h1, s1[2 * k + 1]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { red[0][pl * C + 8 * cg + k] = s0[k]; red[1][pl * C + 8 * cg + k] = s1[k]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t0 = 0.f, t1 = 0.f;
    for (int i = 0; i < PL; ++i) { t0 += red[0][i * C + c]; t1 += red[1][i * C + c]; }
    float* o = part + ((size_t)c * gridDim.y + blockIdx.y) * 2;
    o[0] = t0;
    o[1] = t1;
  }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_partial_smallc_kernel(TV x, TV dy, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps, int act,
                                    float* __restrict__ part, long long per_split) {
  __shared__ float red0[BN_THREADS / 32][BN_SMALLC];
  __shared__ float red1[BN_THREADS / 32][BN_SMALLC];
  const long long HW = (long long)x.h * x.w;
  const long long P = HW * x.n;
  const long long p_lo = (long long)blockIdx.y * per_split;
  const long long p_hi = p_lo + per_split < P ? p_lo + per_split : P;
  float s0[BN_SMALLC], s1[BN_SMALLC], g[BN_SMALLC], b[BN_SMALLC], m[BN_SMALLC], rs[BN_SMALLC];
#pragma unroll
  for (int c = 0; c < BN_SMALLC; ++c) {
    s0[c] = s1[c] = 0.f;
    const bool ok = c < x.c;
    g[c] = (ok && gamma) ? gamma[c] : 1.f;
    b[c] = (ok && beta) ? beta[c] : 0.f;
    m[c] = ok ? mean[c] : 0.f;
    rs[c] = ok ? rsqrtf(var[c] + eps) : 1.f;
  }
  for (long long p = p_lo + threadIdx.x; p < p_hi; p += BN_THREADS) {
    int n, h, w;
    pix_decode(p, HW, x.w, n, h, w);
#pragma unroll
    for (int c = 0; c < BN_SMALLC; ++c)
      if (c < x.c) {
        const float xhat = (x.ld(x.off(n, c, h, w)) - m[c]) * rs[c];
        const float dz = dy.ld(dy.off(n, c, h, w)) * act_grad(fmaf(g[c], xhat, b[c]), act);
        s0[c] += dz;
        s1[c] = fmaf(dz, xhat, s1[c]);
      }
  }
#pragma unroll
  for (int c = 0; c < BN_SMALLC; ++c) {
    for (int o = 16; o > 0; o >>= 1) {
      s0[c] += __shfl_down_sync(0xffffffffu, s0[c], o);
      s1[c] += __shfl_down_sync(0xffffffffu, s1[c], o);
    }
    if ((threadIdx.x & 31) == 0) { red0[threadIdx.x >> 5][c] = s0[c]; red1[threadIdx.x >> 5][c] = s1[c]; }
  }
  __syncthreads();
  if ((int)threadIdx.x < x.c) {
    double t0 = 0.0, t1 = 0.0;
    for (int i = 0; i < BN_THREADS / 32; ++i) { t0 += red0[i][threadIdx.x]; t1 += red1[i][threadIdx.x]; }
    float* o = part + ((size_t)threadIdx.x * gridDim.y + blockIdx.y) * 2;
    o[0] = (float)t0;
    o[1] = (float)t1;
  }
}

__global__ void __launch_bounds__(BN_FIN_THREADS)
bn_bwd_reduce_final_kernel(const float* __restrict__ part, int splits, int C, float* __restrict__ sum_dz,
                           float* __restrict__ sum_dz_xhat) {
  pdl_wait();
  const int c = blockIdx.x * (BN_FIN_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double t0 = 0.0, t1 = 0.0;
  for (int s = lane; s < splits; s += 32) {
    const float* o = part + ((size_t)c * splits + s) * 2;
    t0 += o[0];
    t1 += o[1];
  }
  t0 = warp_sum_d(t0);
  t1 = warp_sum_d(t1);
  if (lane == 0) {
    sum_dz[c] = (float)t0;
    sum_dz_xhat[c] = (float)t1;
  }
}

int launch_bn_bwd_reduce(const TV& x, const TV& dy, const float* gamma, const float* beta,
                         const float* mean, const float* var, float eps, int act, float* sum_dz,
                         float* sum_dz_xhat, cudaStream_t st) {
  if (x.c == 0) return OFA_OK;
  long long per_split = 0;
  const int splits = bn_splits(x, &per_split, 3);
  float* part = nullptr;
  keep_async_pool_resident();
  OFA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&part), (size_t)splits * x.c * 2 * sizeof(float), st));
  int rc;
  if (x.c <= BN_SMALLC) {
    dim3 grid(1, splits);
    bn_bwd_reduce_partial_smallc_kernel<<<grid, BN_THREADS, 0, st>>>(x, dy, gamma, beta, mean, var, eps, act, part,
                                                                    per_split);
    rc = check_launch("bn_bwd_reduce_partial_smallc_kernel");
  } else if (bn_vec8_ok(x) && tv_vec8_ok(dy) && dy.dtype == x.dtype) {
    const int V = x.c / 8;
    dim3 grid(1, splits);
    launch_pdl(bn_bwd_reduce_partial_vec8_kernel, dim3(grid), dim3(V * (256 / V)), 0, st, 
        reinterpret_cast<const uint4*>(x.ptr), reinterpret_cast<const uint4*>(dy.ptr), x.dtype == OFA_F16, V,
        (long long)x.n * x.h * x.w, gamma, beta, mean, var, eps, act, part, per_split);
    rc = check_launch("bn_bwd_reduce_partial_vec8_kernel");
  } else if (tv_pair_ok(x) && tv_pair_ok(dy)) {
    dim3 grid((x.c + 2 * BN_CH - 1) / (2 * BN_CH), splits);
    bn_bwd_reduce_partial_nhwc_kernel<<<grid, BN_THREADS, 0, st>>>(x, dy, gamma, beta, mean, var, eps, act, part,
                                                                  per_split);
    rc = check_launch("bn_bwd_reduce_partial_nhwc_kernel");
  } else {
    dim3 grid((x.c + BN_CH - 1) / BN_CH, splits);
    bn_bwd_reduce_partial_kernel<<<grid, BN_THREADS, 0, st>>>(x, dy, gamma, beta, mean, var, eps, act, part, per_split,
                                                             x.sc == 1);
    rc = check_launch("bn_bwd_reduce_partial_kernel");
  }
  if (!rc) {
    launch_pdl(bn_bwd_reduce_final_kernel, dim3((x.c + BN_FIN_THREADS / 32 - 1) / (BN_FIN_THREADS / 32)), dim3(BN_FIN_THREADS), 0, st, part, splits, x.c, sum_dz, sum_dz_xhat);
    rc = check_launch("bn_bwd_reduce_final_kernel");
  }
  cudaFreeAsync(part, st);
  return rc;
}

__global__ void bn_bwd_apply_kernel(TV x, TV dy, TV dx, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps, int act, int training,
                                    const float* __restrict__ sum_dz,
                                    const float* __restrict__ sum_dz_xhat, int c_is_inner) {
  const long long total = (long long)x.n * x.c * x.h * x.w;
  const float invP = 1.f / (float)((long long)x.n * x.h * x.w);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int n, c, h, w;
    long long r = i;
    if (c_is_inner) {
      c = (int)(r % x.c); r /= x.c;
      w = (int)(r % x.w); r /= x.w;
      h = (int)(r % x.h); n = (int)(r / x.h);
    } else {
      w = (int)(r % x.w); r /= x.w;
      h = (int)(r % x.h); r /= x.h;
      c = (int)(r % x.c); n = (int)(r / x.c);
    }
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    const float m = mean ? mean[c] : 0.f, rstd = var ? rsqrtf(var[c] + eps) : 1.f;
    float xhat = (x.ld(x.off(n, c, h, w)) - m) * rstd;
    float z = fmaf(g, xhat, b);
    float dz = dy.ld(dy.off(n, c, h, w)) * act_grad(z, act);
    float v;
    if (training) v = g * rstd * (dz - sum_dz[c] * invP - xhat * sum_dz_xhat[c] * invP);
    else v = g * rstd * dz;
    dx.st(dx.off(n, c, h, w), v);
  }
}

__global__ void bn_bwd_apply_nhwc_kernel(TV x, TV dy, TV dx, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, const float* __restrict__ mean,
                                         const float* __restrict__ var, float eps, int act, int training,
                                         const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat,
                                         unsigned pairs, unsigned cpairs) {
  const float invP = 1.f / (float)((long long)x.n * x.h * x.w);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += gridDim.x * blockDim.x) {
    const int c = 2 * (int)(i % cpairs);
    const long long o = 2ll * i;
    const float2 xv = tv_ld2(x, o), gv = tv_ld2(dy, o);
    float2 out;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int cc = c + k;
      const float g = gamma ? gamma[cc] : 1.f, b = beta ? beta[cc] : 0.f;
      const float m = mean ? mean[cc] : 0.f, rstd = var ? rsqrtf(var[cc] + eps) : 1.f;
      const float xhat = ((k ? xv.y : xv.x) - m) * rstd;
      const float dz = (k ? gv.y : gv.x) * act_grad(fmaf(g, xhat, b), act);
      const float v = training ? g * rstd * (dz - sum_dz[cc] * invP - xhat * sum_dz_xhat[cc] * invP) : g * rstd * dz;
      if (k) out.y = v; else out.x = v;
    }
    tv_st2(dx, o, out);
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_vec8_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ mean, const float* __restrict__ var, float eps, int act,
                         int training, const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat,
                         float invP, int f16, unsigned nvec, unsigned V) {
  pdl_wait();
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = 8 * (int)(gid % V);
  // dx = a * (dz - k1 - xhat * k2),  xhat = x * rstd + nm,  z = g * xhat + b
  float rstd[8], nm[8], g[8], b[8], a[8], k1[8], k2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k;
    g[k] = gamma ? gamma[c] : 1.f;
    b[k] = beta ? beta[c] : 0.f;
    rstd[k] = var ? rsqrtf(var[c] + eps) : 1.f;
    nm[k] = -(mean ? mean[c] : 0.f) * rstd[k];
    a[k] = g[k] * rstd[k];
    k1[k] = training ? sum_dz[c] * invP : 0.f;
    k2[k] = training ? sum_dz_xhat[c] * invP : 0.f;
  }
  for (unsigned i = gid; i < nvec; i += gridDim.x * blockDim.x) {
    const uint4 xv = x[i], gv = dy[i];
    const uint32_t xi[4] = {xv.x, xv.y, xv.z, xv.w}, gi[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 xf = unpack16(xi[k], f16), gf = unpack16(gi[k], f16);
      const float h0 = fmaf(xf.x, rstd[2 * k], nm[2 * k]), h1 = fmaf(xf.y, rstd[2 * k + 1], nm[2 * k + 1]);
      const float d0 = gf.x * act_grad(fmaf(g[2 * k], h0, b[2 * k]), act);
      const float d1 = gf.y * act_grad(fmaf(g[2 * k + 1], h1, b[2 * k + 1]), act);
      o[k] = pack16(a[2 * k] * (d0 - k1[2 * k] - h0 * k2[2 * k]),
                    a[2 * k + 1] * (d1 - k1[2 * k + 1] - h1 * k2[2 * k + 1]), f16);
    }
    dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

int launch_bn_bwd_apply(const TV& x, const TV& dy, const TV& dx, const float* gamma, const float* beta,
                        const float* mean, const float* var, float eps, int act, int training,
                        const float* sum_dz, const float* sum_dz_xhat, cudaStream_t st) {
  long long total = (long long)x.n * x.c * x.h * x.w;
  if (total == 0) return OFA_OK;
  long long blocks = (total + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (total < (1ll << 34) && tv_vec8_ok(x) && tv_vec8_ok(dy) && tv_vec8_ok(dx) && dy.dtype == x.dtype &&
      dx.dtype == x.dtype) {
    const unsigned nvec = (unsigned)(total / 8), V = (unsigned)(x.c / 8);
    launch_pdl(bn_bwd_apply_vec8_kernel, dim3(vec8_blocks(nvec, V, 256, 3)), dim3(256), 0, st, 
        reinterpret_cast<const uint4*>(x.ptr), reinterpret_cast<const uint4*>(dy.ptr), reinterpret_cast<uint4*>(dx.ptr),
        gamma, beta, mean, var, eps, act, training, sum_dz, sum_dz_xhat,
        1.f / (float)((long long)x.n * x.h * x.w), x.dtype == OFA_F16, nvec, V);
    return check_launch("bn_bwd_apply_vec8_kernel");
  }
  if (total < (1ll << 32) && tv_pair_ok(x) && tv_pair_ok(dy) && tv_pair_ok(dx)) {
    bn_bwd_apply_nhwc_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, dy, dx, gamma, beta, mean, var, eps, act, training,
                                                               sum_dz, sum_dz_xhat, (unsigned)(total / 2),
                                                               (unsigned)(x.c / 2));
    return check_launch("bn_bwd_apply_nhwc_kernel");
  }
  bn_bwd_apply_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, dy, dx, gamma, beta, mean, var, eps, act, training,
                                                        sum_dz, sum_dz_xhat, x.sc == 1);
  return check_launch("bn_bwd_apply_kernel");
}

// =================================================================================================
// (a14) depthwise filter gradient: dW[c, tap] = sum_p x[p + tap] * dy[p]
//   grid (C/32, pixel splits); each thread keeps KS*KS accumulators; block reduce; atomicAdd.
// =================================================================================================
template <int KS>
__global__ void __launch_bounds__(BN_THREADS)
dw_bwd_filter_kernel(TV x, TV dy, float* __restrict__ dw, long long pix_per_block, int c_is_inner) {
  __shared__ float red[BN_PL][BN_CH + 1];
  int cl, pl;
  if (c_is_inner) { cl = threadIdx.x % BN_CH; pl = threadIdx.x / BN_CH; }
  else            { pl = threadIdx.x % BN_PL; cl = threadIdx.x / BN_PL; }
  const int c = blockIdx.x * BN_CH + cl;
  const long long HW = (long long)x.h * x.w;
  const long long P = HW * x.n;
  const long long p_begin = (long long)blockIdx.y * pix_per_block;
  long long p_end = p_begin + pix_per_block;
  if (p_end > P) p_end = P;
  constexpr int R = KS / 2;
  float acc[KS * KS];
#pragma unroll
  for (int j = 0; j < KS * KS; ++j) acc[j] = 0.f;
  if (c < x.c)
    for (long long p = p_begin + pl; p < p_end; p += BN_PL) {
      int n, h, w;
      pix_decode(p, HW, x.w, n, h, w);
      float g = dy.ld(dy.off(n, c, h, w));
#pragma unroll
      for (int ky = 0; ky < KS; ++ky) {
        int ih = h + ky - R;
        if (ih < 0 || ih >= x.h) continue;
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          int iw = w + kx - R;
          if (iw < 0 || iw >= x.w) continue;
          acc[ky * KS + kx] = fmaf(x.ld(x.off(n, c, ih, iw)), g, acc[ky * KS + kx]);
        }
      }
    }
#pragma unroll
  for (int j = 0; j < KS * KS; ++j) {
    red[pl][cl] = acc[j];
    __syncthreads();
    if (threadIdx.x < BN_CH) {
      int cc = blockIdx.x * BN_CH + threadIdx.x;
      if (cc < x.c) {
        float t = 0.f;
        for (int i = 0; i < BN_PL; ++i) t += red[i][threadIdx.x];
        atomicAdd(&dw[(size_t)cc * KS * KS + j], t);
      }
    }
    __syncthreads();
  }
}

// dense-NHWC variant: a thread owns (channel pair, filter row ky, row group) and walks along an image row with a
// KS-wide register window of x, so a pixel costs one x load, one dy load and KS paired FMAs (the one-thread-per-
// pixel form above issues KS*KS loads per pixel and keeps KS*KS accumulators live).  The window rotates by static
// register renaming: the walk is unrolled KS steps.  Rows of x are re-read by the KS filter rows out of L1/L2.
template <int KS>
struct DwRows {
  static constexpr int RG = KS == 3 ? 4 : 2;             // row groups per block
  static constexpr int THREADS = 32 * KS * RG;
  static constexpr int BLOCKS_PER_SM = KS == 7 ? 2 : 3;   // what the register count allows without spills
};

template <bool F32> struct PairRaw;                      // a channel pair as loaded: fp32x2, or two 16-bit values packed
template <> struct PairRaw<true> {
  typedef float2 T;
  static __device__ __forceinline__ T zero() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ float2 f2(T v, bool) { return v; }
};
template <> struct PairRaw<false> {
  typedef uint32_t T;
  static __device__ __forceinline__ T zero() { return 0u; }
  static __device__ __forceinline__ float2 f2(T v, bool h16) { return unpack16(v, h16); }
};

template <int KS, bool F32>
__global__ void __launch_bounds__(DwRows<KS>::THREADS, DwRows<KS>::BLOCKS_PER_SM)
dw_bwd_filter_rows_kernel(TV x, TV dy, float* __restrict__ dw, int rows_per_block) {
  pdl_wait();
  typedef typename PairRaw<F32>::T Raw;
  constexpr int R = KS / 2, RG = DwRows<KS>::RG, T = KS * KS;
  __shared__ float red[RG][64][T];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ky = warp % KS, rg = warp / KS;
  const int H = x.h, W = x.w, C = x.c;
  const int c = (blockIdx.x * 32 + lane) * 2;
  const int rows = x.n * H;
  const int r_begin = blockIdx.y * rows_per_block;
  const int r_end = min(r_begin + rows_per_block, rows);
  float2 acc[KS];
#pragma unroll
  for (int j = 0; j < KS; ++j) acc[j] = make_float2(0.f, 0.f);
  if (c < C) {
    const int C2 = C >> 1;                               // pair stride between pixels
    const bool h16 = x.dtype == OFA_F16;                 // (x and dy have the same type on this path)
    const Raw* xp = reinterpret_cast<const Raw*>(x.ptr) + (c >> 1);
    const Raw* gp = reinterpret_cast<const Raw*>(dy.ptr) + (c >> 1);
    for (int r = r_begin + rg; r < r_end; r += RG) {
      const int ih = r % H + ky - R;
      if (ih < 0 || ih >= H) continue;
      const Raw* xr = xp + (r + ky - R) * W * C2;        // x[n, ih, 0, c]
      const Raw* gr = gp + r * W * C2;
      float2 win[KS];                                    // win[j] = x[ih, w0 + j - R] at the top of a pass
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const int iw = j - R;
        Raw v = PairRaw<F32>::zero();
        if (iw >= 0 && iw < W) v = xr[iw * C2];
        win[j] = PairRaw<F32>::f2(v, h16);
      }
      for (int w0 = 0; w0 < W; w0 += KS) {
        // one pass = KS steps along the row; all 2*KS loads are issued before the first FMA needs them
        Raw g[KS], nx[KS];
#pragma unroll
        for (int u = 0; u < KS; ++u) {
          g[u] = PairRaw<F32>::zero();
          nx[u] = PairRaw<F32>::zero();
          if (w0 + u < W) g[u] = gr[(w0 + u) * C2];
          if (w0 + u + 1 + R < W) nx[u] = xr[(w0 + u + 1 + R) * C2];   // the column entering the window after step u
        }
#pragma unroll
        for (int u = 0; u < KS; ++u) {
          const float2 gv = PairRaw<F32>::f2(g[u], h16);  // zero past the end of the row
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) ptx::ffma2(acc[kx], win[(u + kx) % KS], gv);
          win[u] = PairRaw<F32>::f2(nx[u], h16);
        }
      }
    }
  }
#pragma unroll
  for (int kx = 0; kx < KS; ++kx) {
    red[rg][2 * lane][ky * KS + kx] = acc[kx].x;
    red[rg][2 * lane + 1][ky * KS + kx] = acc[kx].y;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * T; i += DwRows<KS>::THREADS) {
    const int cc = blockIdx.x * 64 + i / T;
    if (cc >= C) break;
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < RG; ++g) t += (&red[g][0][0])[i];
    atomicAdd(&dw[(size_t)blockIdx.x * 64 * T + i], t);
  }
}

template <int KS>
static void launch_dw_rows(const TV& x, const TV& dy, float* dw, cudaStream_t st) {
  const int cb = (x.c + 63) / 64;
  const int rows = x.n * x.h;
  int sp = sm_count() * DwRows<KS>::BLOCKS_PER_SM / cb;
  if (sp < 1) sp = 1;
  int rpb = (rows + sp - 1) / sp;
  if (rpb < DwRows<KS>::RG) rpb = DwRows<KS>::RG;
  sp = (rows + rpb - 1) / rpb;
  dim3 grid(cb, (unsigned)sp);
  if (x.dtype == OFA_F32) launch_pdl(dw_bwd_filter_rows_kernel<KS, true>, dim3(grid), dim3(DwRows<KS>::THREADS), 0, st, x, dy, dw, rpb);
  else launch_pdl(dw_bwd_filter_rows_kernel<KS, false>, dim3(grid), dim3(DwRows<KS>::THREADS), 0, st, x, dy, dw, rpb);
}

int launch_dw_bwd_filter(const TV& x, const TV& dy, int ks, float* dw, cudaStream_t st) {
  if (x.c == 0) return OFA_OK;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)x.c * ks * ks, st);
  if (e != cudaSuccess) return fail(OFA_ERR_CUDA, "memset dw: %s", cudaGetErrorString(e));
  long long P = (long long)x.n * x.h * x.w;
  if (P == 0) return OFA_OK;
  if (tv_pair_ok(x) && tv_pair_ok(dy) && x.dtype == dy.dtype && P * x.c < (1ll << 31)) {
    switch (ks) {
      case 3: launch_dw_rows<3>(x, dy, dw, st); break;
      case 5: launch_dw_rows<5>(x, dy, dw, st); break;
      case 7: launch_dw_rows<7>(x, dy, dw, st); break;
      default: return fail(OFA_ERR_UNSUPPORTED, "depthwise kernel size %d", ks);
    }
    return check_launch("dw_bwd_filter_rows_kernel");
  }
  int cb = (x.c + BN_CH - 1) / BN_CH;
  long long splits = (long long)sm_count() * 4 / cb;
  if (splits < 1) splits = 1;
  long long ppb = (P + splits - 1) / splits;
  if (ppb < 256) ppb = 256;
  splits = (P + ppb - 1) / ppb;
  dim3 grid(cb, (unsigned)splits);
  int ci = x.sc == 1;
  switch (ks) {
    case 3: dw_bwd_filter_kernel<3><<<grid, BN_THREADS, 0, st>>>(x, dy, dw, ppb, ci); break;
    case 5: dw_bwd_filter_kernel<5><<<grid, BN_THREADS, 0, st>>>(x, dy, dw, ppb, ci); break;
    case 7: dw_bwd_filter_kernel<7><<<grid, BN_THREADS, 0, st>>>(x, dy, dw, ppb, ci); break;
    default: return fail(OFA_ERR_UNSUPPORTED, "depthwise kernel size %d", ks);
  }
  return check_launch("dw_bwd_filter_kernel");
}

// chain rule through the filter transform (dynamic_op.py:46-71), one thread per channel, atomics
// for the shared matrices.
// no transform in play (ks == kmax, or KERNEL_TRANSFORM_MODE off): the gradient lands on the centre crop; one thread
// per (channel, tap) so that consecutive threads touch consecutive addresses
__global__ void active_filter_bwd_crop_kernel(int kmax, int ks, int C, const float* __restrict__ dwa,
                                              float* __restrict__ dw7) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * ks * ks) return;
  const int c = i / (ks * ks), t = i - c * ks * ks;
  const int y = t / ks, x = t - y * ks, s = kmax / 2 - ks / 2;
  dw7[(size_t)c * kmax * kmax + (y + s) * kmax + (x + s)] += dwa[i];
}

// transform path: a block of 64 threads owns 64 channels (thread = channel).  The per-channel chain (recompute the
// 5x5, push the gradient back through the 5->3 and 7->5 matrices) is private to the thread; the gradients of the
// SHARED matrices are sums over channels, reduced inside the block through shared memory -- each thread sums a few
// matrix entries over the block's 64 channels -- so a matrix entry receives one atomicAdd per block instead of one
// per channel (240 k atomics on 625 addresses took ~20 us per layer).
constexpr int AFB_CH = 64;
__global__ void __launch_bounds__(AFB_CH)
active_filter_bwd_kernel(const float* __restrict__ w7, int kmax, const float* __restrict__ m75,
                         const float* __restrict__ m53, int ks, int C, const float* __restrict__ dwa,
                         float* __restrict__ dw7, float* __restrict__ dm75, float* __restrict__ dm53) {
  pdl_wait();
  __shared__ float s_src[25][AFB_CH + 1];     // the values the current matrix multiplied (per channel)
  __shared__ float s_g[25][AFB_CH + 1];       // the gradient of that matrix product's output (per channel)
  const int c = blockIdx.x * AFB_CH + threadIdx.x;
  const bool ok = c < C;
  const int nch = min(AFB_CH, C - blockIdx.x * AFB_CH);
  const float* w = w7 + (size_t)(ok ? c : 0) * kmax * kmax;
  float* dw = dw7 + (size_t)(ok ? c : 0) * kmax * kmax;
  const float* g = dwa + (size_t)(ok ? c : 0) * ks * ks;
  const bool step75 = (kmax == 7 && m75 != nullptr);
  // forward recompute of the intermediate 5x5 (when the 7->5 step exists); cur = what the ->3 step reads
  float k5[25], gcur[25];
  const int kc = step75 ? 5 : kmax;           // kc in {5, 7}; with kc == 7 only ks == 3 reaches this kernel (7->3)
  if (step75 && ok) {
    for (int j = 0; j < 25; ++j) {
      float acc = 0.f;
      for (int i = 0; i < 25; ++i) acc = fmaf(w[(i / 5 + 1) * 7 + (i % 5 + 1)], m75[j * 25 + i], acc);
      k5[j] = acc;
    }
  }
  for (int j = 0; j < 25; ++j) gcur[j] = 0.f;
  float g7[49];                               // gradient w.r.t. the 7x7 when the ->3 step reads it directly
  if (!step75) for (int j = 0; j < 49; ++j) g7[j] = 0.f;
  if (ks == kc) {                             // ks == 5 through the 7->5 matrix
    if (ok) for (int j = 0; j < 25; ++j) gcur[j] = g[j];
  } else {                                    // ks == 3: through the ->3 matrix
    const int s = kc / 2 - 1;
    for (int j = 0; j < 9; ++j) {
      const int src = (j / 3 + s) * kc + (j % 3 + s);
      s_src[j][threadIdx.x] = ok ? (step75 ? k5[src] : w[src]) : 0.f;
      s_g[j][threadIdx.x] = ok ? g[j] : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 81; e += AFB_CH) {        // dm53[j][i] = sum_c g[c][j] * cur[c][src(i)]
      const int j = e / 9, i = e - j * 9;
      float acc = 0.f;
      for (int q = 0; q < nch; ++q) acc = fmaf(s_g[j][q], s_src[i][q], acc);
      atomicAdd(&dm53[e], acc);
    }
    __syncthreads();
    if (ok)
      for (int j = 0; j < 9; ++j)
        for (int i = 0; i < 9; ++i) {
          const int src = (i / 3 + s) * kc + (i % 3 + s);
          const float v = g[j] * m53[j * 9 + i];
          if (step75) gcur[src] += v; else g7[src] += v;
        }
  }
  if (step75) {
    for (int j = 0; j < 25; ++j) {
      s_src[j][threadIdx.x] = ok ? w[(j / 5 + 1) * 7 + (j % 5 + 1)] : 0.f;
      s_g[j][threadIdx.x] = ok ? gcur[j] : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 625; e += AFB_CH) {       // dm75[j][i] = sum_c g5[c][j] * w[c][crop(i)]
      const int j = e / 25, i = e - j * 25;
      float acc = 0.f;
      for (int q = 0; q < nch; ++q) acc = fmaf(s_g[j][q], s_src[i][q], acc);
      atomicAdd(&dm75[e], acc);
    }
    if (ok)
      for (int i = 0; i < 25; ++i) {
        float acc = 0.f;
        for (int j = 0; j < 25; ++j) acc = fmaf(gcur[j], m75[j * 25 + i], acc);
        dw[(i / 5 + 1) * 7 + (i % 5 + 1)] += acc;
      }
  } else if (ok) {
    for (int j = 0; j < kc * kc; ++j) dw[j] += g7[j];
  }
}

int launch_active_filter_bwd(const float* w7, int kmax, const float* m75, const float* m53,
                             int transform_on, int ks, int C, const float* dwa, float* dw7, float* dm75,
                             float* dm53, cudaStream_t st) {
  if (C == 0) return OFA_OK;
  if (!transform_on || ks == kmax) {
    const int total = C * ks * ks;
    launch_pdl(active_filter_bwd_crop_kernel, dim3((total + 255) / 256), dim3(256), 0, st, kmax, ks, C, dwa, dw7);
    return check_launch("active_filter_bwd_crop_kernel");
  }
  launch_pdl(active_filter_bwd_kernel, dim3((C + AFB_CH - 1) / AFB_CH), dim3(AFB_CH), 0, st, w7, kmax, m75, m53, ks, C, dwa, dw7, dm75,
                                                                       dm53);
  return check_launch("active_filter_bwd_kernel");
}

// =================================================================================================
// (a14) dense conv weight gradient: dW[o, i, ky, kx] += sum_p dy[p, o] * x[p + tap, i]
//   GEMM with the reduction over pixels: tile 64 outputs x 64 (tap, ci) columns, split over pixels
//   in grid.z, atomicAdd into the strided slice.
// =================================================================================================
__global__ void __launch_bounds__(CV_THREADS)
conv_bwd_weight_kernel(TV x, TV dy, float* __restrict__ dw, long long w_so, long long w_si,
                       long long w_sh, long long w_sw, int cin, int cout, int ks,
                       long long pix_per_block) {
  __shared__ float sA[CV_TK][CV_TO + 4];  // dy  [pixel kk][o]
  __shared__ float sB[CV_TK][CV_TP + 4];  // x   [pixel kk][col]
  const int H = x.h, W = x.w;
  const long long HW = (long long)H * W;
  const long long P = HW * x.n;
  const int o0 = blockIdx.x * CV_TO;
  const int col0 = blockIdx.y * CV_TP;
  const int ncol = ks * ks * cin;
  const int R = ks / 2;
  const long long p_begin = (long long)blockIdx.z * pix_per_block;
  long long p_end = p_begin + pix_per_block;
  if (p_end > P) p_end = P;
  const int tid = threadIdx.x;
  const int to = (tid % 16) * 4;
  const int tc = (tid / 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loaders: thread -> (e = tid % 64 (o or col), kk = tid / 64 + 4*r)
  const int l_e = tid % 64;
  const int l_k = tid / 64;
  // column decode for the x loader
  const int col = col0 + l_e;
  const bool col_ok = col < ncol;
  int ctap = col_ok ? col / cin : 0;
  const int cci = col_ok ? col - ctap * cin : 0;
  const int cky = ctap / ks, ckx = ctap - cky * ks;

  for (long long pb = p_begin; pb < p_end; pb += CV_TK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int kk = l_k + 4 * r;
      long long p = pb + kk;
      float va = 0.f, vb = 0.f;
      if (p < p_end) {
        int n, h, w;
        pix_decode(p, HW, W, n, h, w);
        int o = o0 + l_e;
        if (o < cout) va = dy.ld(dy.off(n, o, h, w));
        int ih = h + cky - R, iw = w + ckx - R;
        if (col_ok && ih >= 0 && ih < H && iw >= 0 && iw < W) vb = x.ld(x.off(n, cci, ih, iw));
      }
      sA[kk][l_e] = va;
      sB[kk][l_e] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CV_TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][to + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tc + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int o = o0 + to + i;
    if (o >= cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int cc = col0 + tc + j;
      if (cc >= ncol) continue;
      int tap = cc / cin, ci = cc - tap * cin;
      int ky = tap / ks, kx = tap - ky * ks;
      atomicAdd(&dw[o * w_so + ci * w_si + ky * w_sh + kx * w_sw], acc[i][j]);
    }
  }
}

int launch_conv_bwd_weight(const TV& x, const TV& dy, float* dw, long long w_so, long long w_si,
                           long long w_sh, long long w_sw, int cin, int cout, int ks, cudaStream_t st) {
  long long P = (long long)x.n * x.h * x.w;
  if (P == 0 || cin == 0 || cout == 0) return OFA_OK;
  int gx = (cout + CV_TO - 1) / CV_TO;
  int gy = (ks * ks * cin + CV_TP - 1) / CV_TP;
  long long splits = (long long)sm_count() * 2 / ((long long)gx * gy);
  if (splits < 1) splits = 1;
  long long ppb = (P + splits - 1) / splits;
  ppb = (ppb + CV_TK - 1) / CV_TK * CV_TK;
  if (ppb < 256) ppb = 256;
  splits = (P + ppb - 1) / ppb;
  if (splits > 65535) return fail(OFA_ERR_UNSUPPORTED, "conv_bwd_weight: too many pixel splits");
  dim3 grid(gx, gy, (unsigned)splits);
  conv_bwd_weight_kernel<<<grid, CV_THREADS, 0, st>>>(x, dy, dw, w_so, w_si, w_sh, w_sw, cin, cout, ks, ppb);
  return check_launch("conv_bwd_weight_kernel");
}

// =================================================================================================
// (§8f-1) evaluation metric on the device: per-image sum of squared differences of the BT.601 luma of the
// uint8-rounded images — psnr(rgb2y(tensor2img_np(a)), rgb2y(tensor2img_np(b))), sr_run_manager.py:364,567-597.
// Integer result (exact): clamp -> x255 in fp32 -> round half to even -> uint8; Y = rint((65.481 R + 128.553 G
// + 24.966 B) / 255 + 16) in fp64.  grid = (pixel blocks, images); one 64-bit atomic per block.
// =================================================================================================
__device__ __forceinline__ int luma_u8(const TV& t, int n, int h, int w) {
  double acc = 0.0;
  const double k[3] = {65.481, 128.553, 24.966};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = fminf(fmaxf(t.ld(t.off(n, c, h, w)), 0.f), 1.f);
    acc += k[c] * (double)rintf(v * 255.0f);
  }
  return (int)rint(acc / 255.0 + 16.0);
}

__global__ void psnr_y_sse_kernel(TV a, TV b, unsigned long long* __restrict__ sse) {
  __shared__ unsigned long long red[8];
  const int n = blockIdx.y;
  const int HW = a.h * a.w;
  unsigned long long s = 0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const int h = p / a.w, w = p - h * a.w;
    const int d = luma_u8(a, n, h, w) - luma_u8(b, n, h, w);
    s += (unsigned long long)(d * d);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(&sse[n], t);
  }
}

int launch_psnr_y_sse(const TV& a, const TV& b, long long* sse, cudaStream_t st) {
  if (a.n == 0) return OFA_OK;
  cudaError_t e = cudaMemsetAsync(sse, 0, sizeof(long long) * (size_t)a.n, st);
  if (e != cudaSuccess) return fail(OFA_ERR_CUDA, "memset sse: %s", cudaGetErrorString(e));
  const long long HW = (long long)a.h * a.w;
  if (HW == 0) return OFA_OK;
  long long bx = (HW + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (bx > cap) bx = cap;
  dim3 grid((unsigned)bx, (unsigned)a.n);
  psnr_y_sse_kernel<<<grid, 256, 0, st>>>(a, b, reinterpret_cast<unsigned long long*>(sse));
  return check_launch("psnr_y_sse_kernel");
}

// =================================================================================================
// (§8f-3) multi-tensor Adam: torch.optim.Adam as sr_run_manager.py:115-133 builds it (L2 weight decay per
// parameter group: the 'bn#bias' keys get 0).  One launch updates every ACTIVE parameter tensor; tensors whose
// gradient pointer is NULL (blocks outside the sampled sub-network) are skipped entirely — their step counters
// and moments do not move, exactly like optimizer.step() skipping p.grad is None.
// =================================================================================================
__global__ void adam_bump_steps_kernel(const float* const* __restrict__ grads, int* __restrict__ steps, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && grads[t] != nullptr) steps[t] += 1;
}

constexpr int ADAM_CHUNK = 256 * 8;

__global__ void __launch_bounds__(256)
adam_step_kernel(const OfaAdamTensor* __restrict__ table, const int2* __restrict__ chunks,
                 const float* const* __restrict__ grads, const int* __restrict__ steps, float lr, float beta1,
                 float beta2, float eps) {
  const int2 ck = chunks[blockIdx.x];
  const float* __restrict__ g = grads[ck.x];
  if (g == nullptr) return;
  const OfaAdamTensor T = table[ck.x];
  const int step = steps[ck.x];
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  const float step_size = lr / bc1;
  const long long base = (long long)ck.y * ADAM_CHUNK;
#pragma unroll
  for (int k = 0; k < ADAM_CHUNK / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < T.numel) {
      const float p = T.p[i];
      const float gi = fmaf(T.weight_decay, p, g[i]);
      const float m = T.m[i] = beta1 * T.m[i] + (1.f - beta1) * gi;
      const float v = T.v[i] = beta2 * T.v[i] + (1.f - beta2) * gi * gi;
      const float denom = sqrtf(v) / bc2_sqrt + eps;
      T.p[i] = p - step_size * (m / denom);
    }
  }
}

int launch_adam_step(const OfaAdamTensor* table, const int* chunks, int n_tensors, int n_chunks,
                     const float* const* grads, int* steps, float lr, float beta1, float beta2, float eps,
                     cudaStream_t st) {
  if (n_tensors == 0 || n_chunks == 0) return OFA_OK;
  adam_bump_steps_kernel<<<(n_tensors + 127) / 128, 128, 0, st>>>(grads, steps, n_tensors);
  int rc = check_launch("adam_bump_steps_kernel");
  if (rc) return rc;
  adam_step_kernel<<<n_chunks, 256, 0, st>>>(table, reinterpret_cast<const int2*>(chunks), grads, steps, lr, beta1, beta2,
                                             eps);
  return check_launch("adam_step_kernel");
}

}  // namespace ofa
