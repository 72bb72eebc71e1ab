// The stems on the tensor cores: k x k conv (k = 3 | 5) from cin <= 4 channels (the RGB image, fp32, any layout) to 64
// channels, NHWC 16-bit out, folded BatchNorm + activation -- first_conv of OFAMobileNetS4 (ofa_mbs4.py:36-39, ConvLayer
// layers.py:135-147), and through swapped / negative weight strides the data gradient of the 64 -> 3 output conv.
//
// conv_stem_kernel (conv_thin.cu) does this on the CUDA cores: 75 x 64 FMAs per pixel, 0.21 ms for the 960 x 540 frame
// at 35 % of the fp32 pipe.  As a GEMM the layer is D[pixel, 64] = A[pixel, K] * W[64, K]^T with K = k*k*4 (the channel
// dimension padded 3 -> 4): 100 -> 112 for k = 5, 36 -> 48 for k = 3 -- seven (three) M128 x N64 x K16 tcgen05.mma per
// 128 pixels.  TMA cannot build that A operand (im2col of a 3-channel planar fp32 image), so the CTA's threads do: the
// tile's halo is staged in shared memory as [y][x] float4 pixels, a thread owns one pixel = one row of A and writes its
// K values as 16-byte chunks (two taps x four channels each: one LDS.128 + two packs per tap) straight into the
// 128-byte-swizzled K-major layout the UMMA descriptor reads.  ~100 instructions per pixel instead of ~2 400 FMAs.
//   * CTA tile = 16 x 16 pixels = two M-tiles (rows of A = the 128 pixels of an 8 x 16 half); 256 threads build, one
//     thread issues the 2 x 7 MMAs, all eight warps run the epilogue (TMEM lane quarter = warp % 4);
//   * the 64 x K weight slice is converted once per persistent CTA from the fp32 master (any strides);
//   * two CTAs per SM overlap each other's build / MMA / epilogue phases.
#include "ofa_common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

#include <stdlib.h>
#include <string.h>

namespace ofa {
namespace {

constexpr int SC_TH = 16, SC_TW = 16;               // CTA tile (pixels)
constexpr int SC_THREADS = 256;
constexpr int SC_MAXK = 5;
constexpr int SC_HALO = SC_TW + SC_MAXK - 1;        // 20
constexpr int SC_ATOM = 128 * 128;                  // one swizzle atom of A: 128 rows x 128 bytes (64 K-elements)
constexpr int SC_A_BYTES = 2 * 2 * SC_ATOM;         // two M-tiles x two K atoms
constexpr int SC_B_BYTES = 2 * 64 * 128;            // two K atoms x 64 rows
constexpr int SC_IN_BYTES = SC_HALO * SC_HALO * 16; // float4 per halo pixel
constexpr int SC_SMEM = 1024 + SC_A_BYTES + SC_B_BYTES + SC_IN_BYTES + 2 * 64 * 4 + 64;

struct StemTcParams {
  TV x, y;
  const float* w; long long w_so, w_si, w_sh, w_sw;
  int cin, ks, f16;
  Epi epi;
  int tiles_x, tiles_y, num_tiles;
};

__device__ __forceinline__ uint32_t swz(int row, int chunk) {   // byte offset of 16-byte chunk `chunk` of `row` in an atom
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int KS>
__global__ void __launch_bounds__(SC_THREADS, 2)
conv_stem_tc_kernel(const StemTcParams p) {
  pdl_wait();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                   // [mtile][atom][128 rows][128 B]
  uint8_t* sB = sA + SC_A_BYTES;                        // [atom][64 rows][128 B]
  float4* s_in = reinterpret_cast<float4*>(sB + SC_B_BYTES);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_in) + SC_IN_BYTES);
  float* s_shift = s_scale + 64;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_shift + 64);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int ks = KS, R = ks >> 1;                   // compile-time: the im2col offsets below fold to constants
  const int cin = p.cin, f16 = p.f16;
  constexpr int taps = ks * ks;
  constexpr int nchunks = (taps + 1) >> 1;              // 16-byte chunks of real data per row (two taps each)
  constexpr int ksteps = (taps * 4 + 15) >> 4;          // MMAs per M-tile: 7 (k = 5) or 3 (k = 3)
  constexpr int kchunks = ksteps * 2;                   // chunks the MMAs read (the tail beyond nchunks is zero)
  const int H = p.x.h, W = p.x.w;

  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (warp == 1) { ptx::tmem_alloc(tmem_ptr, 128); ptx::tmem_relinquish(); }
  if (tid < 64) {
    float sc, sh;
    epi_scale_shift(p.epi, tid, sc, sh);
    s_scale[tid] = sc; s_shift[tid] = sh;
  }
  // weights: B[co][k], k = tap * 4 + ci, as swizzled K-major 16-byte chunks (two taps x four channels)
  for (int i = tid; i < 64 * kchunks; i += SC_THREADS) {
    const int co = i / kchunks, ch = i - co * kchunks;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int tap = ch * 2 + (e >> 2), ci = e & 3;
      v[e] = 0.f;
      if (tap < taps && ci < cin) {
        const int ky = tap / ks, kx = tap - ky * ks;
        v[e] = p.w[co * p.w_so + ci * p.w_si + ky * p.w_sh + kx * p.w_sw];
      }
    }
    *reinterpret_cast<uint4*>(sB + (ch >> 3) * (64 * 128) + swz(co, ch & 7)) =
        make_uint4(pack16(v[0], v[1], f16), pack16(v[2], v[3], f16), pack16(v[4], v[5], f16), pack16(v[6], v[7], f16));
  }
  // zero the chunks of A that no pixel writes (K padding): once, they stay zero
  for (int i = tid; i < 2 * 128 * (kchunks - nchunks); i += SC_THREADS) {
    const int ch = nchunks + i % (kchunks - nchunks), r = (i / (kchunks - nchunks)) & 127, m = i / ((kchunks - nchunks) * 128);
    *reinterpret_cast<uint4*>(sA + m * (2 * SC_ATOM) + (ch >> 3) * SC_ATOM + swz(r, ch & 7)) = make_uint4(0u, 0u, 0u, 0u);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t idesc = ptx::umma_idesc_f16(128, 64, f16 ? 0 : 1, f16 ? 0 : 1, 0, 0);
  const int per_img = p.tiles_x * p.tiles_y;
  constexpr int hh = SC_TH + ks - 1, hw = SC_TW + ks - 1;
  uint32_t phase = 0;

  // thread -> pixel (py, px) of the tile = row (py & 7) * 16 + px of M-tile py >> 3
  const int py = tid >> 4, px = tid & 15;
  const int mt = py >> 3, arow = (py & 7) * 16 + px;
  uint8_t* a_mine = sA + mt * (2 * SC_ATOM);

  for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
    const int n = t / per_img, r = t - n * per_img;
    const int y0 = (r / p.tiles_x) * SC_TH, x0 = (r % p.tiles_x) * SC_TW;
    // ---- stage the halo: [y][x] float4 (channels beyond cin are zero), zero outside the image = the conv's padding
    for (int i = tid; i < hh * hw; i += SC_THREADS) {
      const int xx = i % hw, yy = i / hw;
      const int iy = y0 + yy - R, ix = x0 + xx - R;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        v.x = p.x.ld(p.x.off(n, 0, iy, ix));
        if (cin > 1) v.y = p.x.ld(p.x.off(n, 1, iy, ix));
        if (cin > 2) v.z = p.x.ld(p.x.off(n, 2, iy, ix));
        if (cin > 3) v.w = p.x.ld(p.x.off(n, 3, iy, ix));
      }
      s_in[yy * SC_HALO + xx] = v;
    }
    __syncthreads();
    // ---- im2col: this thread's row of A
#pragma unroll
    for (int ch = 0; ch < nchunks; ++ch) {
      const int t0 = ch * 2, t1 = t0 + 1;
      const int ky0 = t0 / ks, kx0 = t0 - ky0 * ks;
      const float4 a = s_in[(py + ky0) * SC_HALO + px + kx0];
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t1 < taps) { const int ky1 = t1 / ks, kx1 = t1 - ky1 * ks; b = s_in[(py + ky1) * SC_HALO + px + kx1]; }
      *reinterpret_cast<uint4*>(a_mine + (ch >> 3) * SC_ATOM + swz(arow, ch & 7)) =
          make_uint4(pack16(a.x, a.y, f16), pack16(a.z, a.w, f16), pack16(b.x, b.y, f16), pack16(b.z, b.w, f16));
    }
    ptx::fence_proxy_async();            // generic-proxy writes of A -> visible to the tensor core's async proxy
    __syncthreads();
    // ---- MMAs: one thread, both M-tiles
    if (tid == 0) {
      ptx::tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t a_addr = ptx::smem_u32(sA) + (uint32_t)(m * (2 * SC_ATOM) + (k >> 2) * SC_ATOM);
          const uint32_t b_addr = ptx::smem_u32(sB) + (uint32_t)((k >> 2) * (64 * 128));
          const uint64_t da = ptx::umma_desc_sw128(a_addr, 1024) + (uint64_t)((k & 3) * 2);
          const uint64_t db = ptx::umma_desc_sw128(b_addr, 1024) + (uint64_t)((k & 3) * 2);
          ptx::umma_bf16(tmem_base + (uint32_t)(m * 64), da, db, idesc, k > 0 ? 1u : 0u);
        }
      }
      ptx::umma_commit(bar);
    }
    // ---- epilogue: warp w reads lanes [32 * (w % 4), +32) of M-tile w / 4
    ptx::mbar_wait(bar, phase);
    phase ^= 1;
    ptx::tc_fence_after();
    {
      const int em = warp >> 2, quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const int Y = y0 + em * 8 + (row >> 4), X = x0 + (row & 15);
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(em * 64);
      const bool ok = Y < H && X < W;
      uint16_t* yp = reinterpret_cast<uint16_t*>(p.y.ptr) + p.y.off(n, 0, ok ? Y : 0, ok ? X : 0);
      // two halves of 32 columns: affine from 16-byte shared-memory broadcasts, ONE uniform activation branch per half
      // (a per-element activation switch + residual branch made this loop body 4 500 instructions long)
      const int act = p.epi.act;
#pragma unroll
      for (int h0 = 0; h0 < 64; h0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld16(t_addr + (uint32_t)h0, v);
        ptx::tmem_ld16(t_addr + (uint32_t)(h0 + 16), v + 16);
        ptx::tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + h0 + 4 * q);
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + h0 + 4 * q);
          f[4 * q] = fmaf(__uint_as_float(v[4 * q]), sc.x, sh.x);
          f[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), sc.y, sh.y);
          f[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), sc.z, sh.z);
          f[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), sc.w, sh.w);
        }
        if (act == OFA_ACT_RELU6) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fminf(fmaxf(f[i], 0.f), 6.f);
        } else if (act == OFA_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
        } else if (act != OFA_ACT_NONE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = apply_act(f[i], act);
        }
        if (ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(yp + h0 + 8 * q) =
                make_uint4(pack16(f[8 * q], f[8 * q + 1], f16), pack16(f[8 * q + 2], f[8 * q + 3], f16),
                           pack16(f[8 * q + 4], f[8 * q + 5], f16), pack16(f[8 * q + 6], f[8 * q + 7], f16));
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();                     // accumulators drained, A and the halo may be overwritten
  }
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 128); }
}

}  // namespace

bool conv_stem_tc_supported(const OfaConvArgs* a) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("OFA_STEM_TC"); off = (e && e[0] == '0') ? 1 : 0; }
  if (off) return false;
  if (a->flip || a->store != OFA_STORE_PLAIN || a->epi.residual) return false;
  if (a->cin < 1 || a->cin > 4 || a->cout != 64) return false;
  if (a->ks != 3 && a->ks != 5) return false;
  if (!a->w || a->x.n <= 0 || a->x.h <= 0 || a->x.w <= 0) return false;
  if (!is_16bit(a->y.dtype) || !is_nhwc_dense(&a->y) || (reinterpret_cast<uintptr_t>(a->y.ptr) & 15)) return false;
  if ((long long)a->x.n * ((a->x.h + SC_TH - 1) / SC_TH) * ((a->x.w + SC_TW - 1) / SC_TW) >= (1ll << 31)) return false;
  return true;
}

int launch_conv_stem_tc(const OfaConvArgs* a, cudaStream_t st) {
  StemTcParams p;
  memset(&p, 0, sizeof(p));
  p.x = make_tv(&a->x); p.y = make_tv(&a->y);
  p.w = a->w; p.w_so = a->w_so; p.w_si = a->w_si; p.w_sh = a->w_sh; p.w_sw = a->w_sw;
  p.cin = a->cin; p.ks = a->ks; p.f16 = a->y.dtype == OFA_F16 ? 1 : 0;
  p.epi = make_epi(&a->epi);
  p.tiles_x = (a->x.w + SC_TW - 1) / SC_TW;
  p.tiles_y = (a->x.h + SC_TH - 1) / SC_TH;
  p.num_tiles = a->x.n * p.tiles_x * p.tiles_y;
  static unsigned char attr_done[64] = {0};
  if (once_per_device(attr_done))
  {
    OFA_CUDA(cudaFuncSetAttribute(conv_stem_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
    OFA_CUDA(cudaFuncSetAttribute(conv_stem_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
  }
  int blocks = 2 * sm_count();
  if (blocks > p.num_tiles) blocks = p.num_tiles;
  if (p.ks == 3) launch_pdl(conv_stem_tc_kernel<3>, dim3(blocks), dim3(SC_THREADS), (size_t)SC_SMEM, st, p);
  else launch_pdl(conv_stem_tc_kernel<5>, dim3(blocks), dim3(SC_THREADS), (size_t)SC_SMEM, st, p);
  return check_launch("conv_stem_tc_kernel");
}

}  // namespace ofa
