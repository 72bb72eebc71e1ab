// Kernel-size-elastic depthwise convolution, NHWC bf16, with the folded BatchNorm + ReLU6 epilogue
// (dynamic_op.py:46-84 + dynamic_layers.py:48-57 in inference mode).
//
//   * a CTA owns a TH x TW pixel tile of one 64-channel chunk; the (TH+ks-1) x (TW+ks-1) x 64 halo
//     box arrives with ONE TMA load whose out-of-bounds zero fill is the conv's zero padding;
//   * the active filters (7x7 weights through the learned 7->5 / 5->3 matrices, rotated for the data gradient)
//     are derived ONCE per launch by active_filter_kernel (simt_kernels.cu, chunked layout); a CTA fetches its chunk's ks*ks x 64 floats with one bulk
//     copy on the same mbarrier as the halo (round 1 re-derived them in every tile's CTA: dependent global loads that
//     held 42 % of the kernel's stall samples and made ks = 3 cost as much as ks = 7);
//   * lane = channel pair, so every shared-memory access of a warp is one conflict-free 128-byte
//     pixel row; each thread slides a window over 16 output pixels per pass so one smem word feeds
//     up to ks taps; math is packed fp32x2 FMA (fma.rn.f32x2), fp32 accumulation.
#include "ofa_common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

#include <string.h>

namespace ofa {
namespace {

constexpr int CH = 64;
// tile = TH output rows (one warp each) x TW output columns; {8, 32} for frames, {12, 24} when the image divides into
// such tiles (the 24 x 24 LR training patches: 2 tiles per image instead of 3 with a quarter of each wasted, and 24
// instead of 16 resident warps per SM at the same two CTAs)
// output pixels per pass: acc[RUN] + in[RUN + KS - 1] channel pairs must stay in registers (16 spilled at ks >= 5)
template <int KS, int TW> struct DwRun {
  static constexpr int RUN = (KS == 3 && TW % 16 == 0) ? 16 : 8;
  static_assert(TW % RUN == 0, "passes must tile the row exactly: a partial pass would read past the halo row");
};

struct DwParams {
  int N, H, W, C;
  int kmax, transform_on, flip;
  const float* w7;
  const float* m75;
  const float* m53;
  const float* filt;   // [C / 64][KS * KS][64] fp32 active filters (already rotated when flip), from launch_active_filter_chunked
  const float* gamma; const float* beta; const float* mean; const float* var; float eps;
  int act;
  uint16_t* y;
  int f16;   // storage format of x and y: 1 = fp16, 0 = bf16
  int tiles_w, tiles_h;
};

template <int KS, int TH, int TW>
struct DwSmem {
  static constexpr int HALO_H = TH + KS - 1;
  static constexpr int HALO_W = TW + KS - 1;
  static constexpr int TILE_BYTES = HALO_H * HALO_W * CH * 2;
  static constexpr int FILT_OFF = (TILE_BYTES + 127) / 128 * 128;
  static constexpr int FILT_BYTES = KS * KS * CH * 4;
  static constexpr int SS_OFF = FILT_OFF + FILT_BYTES;
  static constexpr int BAR_OFF = SS_OFF + 2 * CH * 4;
  static constexpr int TOTAL = BAR_OFF + 16 + 128;  // + alignment slack
};

__device__ __forceinline__ float2 bf2_unpack(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

template <int KS, int TH, int TW>
__global__ void __launch_bounds__(32 * TH, TH == 12 ? 3 : 2)
dw_fast_kernel(const __grid_constant__ CUtensorMap tmap_x, const DwParams p) {
  pdl_wait();
  using L = DwSmem<KS, TH, TW>;
  constexpr int THREADS = 32 * TH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u);   // keeps the shared address space
  const uint32_t* tile = reinterpret_cast<const uint32_t*>(smem);           // [HALO_H][HALO_W][32] words
  float* filt = reinterpret_cast<float*>(smem + L::FILT_OFF);               // [KS*KS][CH]
  float* s_scale = reinterpret_cast<float*>(smem + L::SS_OFF);
  float* s_shift = s_scale + CH;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);

  const int tid = threadIdx.x;
  const int c0 = blockIdx.y * CH;
  const int t = blockIdx.x;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int n = t / tiles_per_img;
  const int r = t - n * tiles_per_img;
  const int h0 = (r / p.tiles_w) * TH, w0 = (r % p.tiles_w) * TW;
  constexpr int R = KS / 2;

  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)(L::TILE_BYTES + L::FILT_BYTES));
    ptx::tma_load_4d(smem, &tmap_x, bar, c0, w0 - R, h0 - R, n);
    // the chunk's active filters were derived once per launch (active_filter_kernel): one 12.5 KB bulk copy instead of
    // every tile's CTA re-deriving them with dependent global loads (42 % of the warp-stall samples of the old kernel)
    ptx::bulk_load(filt, p.filt + (size_t)blockIdx.y * KS * KS * CH, (uint32_t)L::FILT_BYTES, bar);
  }
  if (tid < CH) {
    int c = c0 + tid;
    float g = p.gamma ? p.gamma[c] : 1.f, b = p.beta ? p.beta[c] : 0.f;
    float m = p.mean ? p.mean[c] : 0.f, rstd = p.var ? rsqrtf(p.var[c] + p.eps) : 1.f;
    s_scale[tid] = g * rstd;
    s_shift[tid] = b - m * g * rstd;
  }
  __syncthreads();
  ptx::mbar_wait(bar, 0);

  // ---- main: warp = output row, lane = channel pair, RUN output pixels per pass -------------------
  const int warp = tid >> 5, lane = tid & 31;
  const int h = h0 + warp;
  const float2 sc = make_float2(s_scale[2 * lane], s_scale[2 * lane + 1]);
  const float2 sh = make_float2(s_shift[2 * lane], s_shift[2 * lane + 1]);
  const float2* filt2 = reinterpret_cast<const float2*>(filt);  // [KS*KS][32]
  constexpr int RUN = DwRun<KS, TW>::RUN;
#pragma unroll 1
  for (int x0 = 0; x0 < TW && w0 + x0 < p.W; x0 += RUN) {
    float2 acc[RUN];
#pragma unroll
    for (int i = 0; i < RUN; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll(KS >= 5 ? 1 : KS)                       // ks >= 5 fully unrolled hoists rows of loads and spills
    for (int ky = 0; ky < KS; ++ky) {
      const uint32_t* row = tile + ((warp + ky) * L::HALO_W + x0) * 32 + lane;
      float2 in[RUN + KS - 1];
#pragma unroll
      for (int i = 0; i < RUN + KS - 1; ++i) in[i] = unpack16(row[i * 32], p.f16);
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const float2 wv = filt2[(ky * KS + kx) * 32 + lane];
#pragma unroll
        for (int i = 0; i < RUN; ++i) ptx::ffma2(acc[i], in[i + kx], wv);
      }
    }
    if (h < p.H) {
      uint16_t* yrow = p.y + (((size_t)n * p.H + h) * p.W) * p.C + c0 + 2 * lane;
#pragma unroll
      for (int i = 0; i < RUN; ++i) {
        int w = w0 + x0 + i;
        if (w < p.W) {
          float a = apply_act(fmaf(acc[i].x, sc.x, sh.x), p.act);
          float b = apply_act(fmaf(acc[i].y, sc.y, sh.y), p.act);
          *reinterpret_cast<uint32_t*>(yrow + (size_t)w * p.C) = pack16(a, b, p.f16);
        }
      }
    }
  }
}

template <int KS, int TH, int TW>
int launch_tile(const OfaTensor4* x, DwParams& p, cudaStream_t st) {
  using L = DwSmem<KS, TH, TW>;
  p.tiles_w = (p.W + TW - 1) / TW;
  p.tiles_h = (p.H + TH - 1) / TH;
  CUtensorMap tm;
  uint64_t dims[4] = {(uint64_t)p.C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
  uint64_t strides[3] = {(uint64_t)p.C * 2, (uint64_t)p.W * p.C * 2, (uint64_t)p.H * p.W * p.C * 2};
  uint32_t box[4] = {CH, (uint32_t)(TW + KS - 1), (uint32_t)(TH + KS - 1), 1};
  int rc = encode_tmap(&tm, p.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x->ptr, dims,
                       strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  static unsigned char attr_done[64] = {0};   // one per template instance
  if (once_per_device(attr_done))
    OFA_CUDA(cudaFuncSetAttribute(dw_fast_kernel<KS, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  dim3 grid((unsigned)(p.N * p.tiles_w * p.tiles_h), p.C / CH);
  launch_pdl(dw_fast_kernel<KS, TH, TW>, dim3(grid), dim3(32 * TH), L::TOTAL, st, tm, p);
  return check_launch("dw_fast_kernel");
}

template <int KS>
int launch_ks(const OfaTensor4* x, DwParams& p, cudaStream_t st) {
  if (p.H % 12 == 0 && p.W % 24 == 0 && (p.H % 8 != 0 || p.W % 32 != 0)) return launch_tile<KS, 12, 24>(x, p, st);
  return launch_tile<KS, 8, 32>(x, p, st);
}

}  // namespace

bool dw_fast_supported(const OfaTensor4* x, const OfaTensor4* y, int ks, const OfaEpilogue* epi) {
  if (!is_16bit(x->dtype) || y->dtype != x->dtype) return false;
  if (!is_nhwc_dense(x) || !is_nhwc_dense(y)) return false;
  if (x->c % CH != 0 || x->c == 0) return false;
  if (ks != 3 && ks != 5 && ks != 7) return false;
  if (epi && epi->residual) return false;
  if ((reinterpret_cast<uintptr_t>(x->ptr) & 15) || (reinterpret_cast<uintptr_t>(y->ptr) & 3)) return false;
  if (x->n <= 0 || x->h <= 0 || x->w <= 0) return false;
  return true;
}

int launch_dw_fast(const OfaTensor4* x, const OfaTensor4* y, const float* w7, int kmax, const float* m75,
                   const float* m53, int transform_on, int ks, int flip, const OfaEpilogue* epi, cudaStream_t st) {
  if (kmax != 7 && transform_on && ks < kmax && kmax != 5)
    return fail(OFA_ERR_UNSUPPORTED, "dw_fast: kmax %d", kmax);
  DwParams p;
  memset(&p, 0, sizeof(p));
  p.N = x->n; p.H = x->h; p.W = x->w; p.C = x->c;
  p.kmax = kmax; p.transform_on = transform_on; p.flip = flip;
  p.w7 = w7; p.m75 = m75; p.m53 = m53;
  if (epi) {
    p.gamma = epi->gamma; p.beta = epi->beta; p.mean = epi->mean; p.var = epi->var; p.eps = epi->eps;
    p.act = epi->act;
  }
  p.y = reinterpret_cast<uint16_t*>(y->ptr);
  p.f16 = x->dtype == OFA_F16 ? 1 : 0;
  // the active (transformed, for the data gradient rotated) filters of all channels, once per launch
  float* filt = nullptr;
  keep_async_pool_resident();
  OFA_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&filt), (size_t)p.C * ks * ks * sizeof(float), st));
  int rc = launch_active_filter_chunked(w7, kmax, m75, m53, transform_on, ks, p.C, flip, filt, st);
  p.filt = filt;
  if (!rc) {
    switch (ks) {
      case 3: rc = launch_ks<3>(x, p, st); break;
      case 5: rc = launch_ks<5>(x, p, st); break;
      default: rc = launch_ks<7>(x, p, st); break;
    }
  }
  cudaFreeAsync(filt, st);
  return rc;
}

}  // namespace ofa
