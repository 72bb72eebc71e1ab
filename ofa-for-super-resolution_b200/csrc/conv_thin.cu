// The two thin ends of the SR nets (layers.py:94-98,135-147 as used at ofa_mbs4.py:22-31,117-133 and
// ofa_mbx4.py:22-60): both are HBM-bound, and the generic implicit GEMM wastes its time elsewhere.
//
//  (1) conv_out_rows_kernel — k x k conv 64 -> cout with ks * cout <= 16 (S4: 5x5 64->3 at 4x resolution,
//      X4: 3x3 64->3), NHWC 16-bit in, any-layout out (the user's NCHW fp32 image).
//      The generic kernel re-reads its activation tile once per filter tap (25x) from shared memory for an
//      N = 16 MMA and is bound by that port.  Here the kx taps are folded into the accumulator columns:
//          P[x][kx*cout + co] = sum_{ky, ci} in[y + ky - R][x][ci] * w[co][ci][ky][kx]     (MMA, N = 16)
//          out[y][x][co]      = sum_kx P[x + kx - R][kx*cout + co]                          (epilogue)
//      so an image row is read from shared memory ks times instead of ks*ks times, and the shift along x
//      is a transposed pass through a small shared-memory buffer in the epilogue.  A CTA walks down a
//      128-pixel-wide strip with a rolling ring of image rows (each row is loaded once per row block),
//      M = 128 consecutive pixels of one row, K = 64 channels.
//
//  (2) conv_stem_kernel — k x k conv cin <= 4 -> 64 (the 5x5 / 3x3 stems), any-layout fp32 in, NHWC 16-bit
//      out: CUDA cores with a register-blocked sliding window (4 pixels x 16 channels per thread), the
//      whole weight tensor in shared memory.  2.5 GFMA at C2 shapes: not worth a tensor-core im2col.
#include "ofa_common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

#include <string.h>

namespace ofa {
namespace {

// ==================================================================================================
// (1) thin-output conv on tcgen05
// ==================================================================================================
constexpr int CO_MPIX = 128;               // pixels of one image row per tile = UMMA M
constexpr int CO_RING = 10;                // image rows resident (16 KiB each)
constexpr int CO_ROW_BYTES = CO_MPIX * 128;
constexpr int CO_RB = 32;                  // output rows per work item
constexpr int CO_ACC = 8;                  // accumulator stages (16 TMEM columns each)
constexpr int CO_PPITCH = CO_MPIX + 4;     // floats per row of the transpose buffer
constexpr int CO_THREADS = 11 * 32;        // TMA, MMA issuer 0, epilogue set 0 (4 warps), epilogue set 1 (4), MMA issuer 1
constexpr int CO_ISSUER1_WARP = 10;

struct ConvOutParams {
  int N, H, W, cout, ks, f16;
  const float* w; long long w_so, w_si, w_sh, w_sw;
  int strips, row_blocks;
  Epi epi;
  TV y;
};

// KS_ / COUT_: compile-time kernel size and output channels (0 = take them from the parameters); OUT_: output format
// known at compile time (OFA_F32 / OFA_U8 planar images; -1 = any, through TV::st)
template <int KS_, int COUT_, int OUT_>
__global__ void __launch_bounds__(CO_THREADS, 1)
conv_out_rows_kernel(const __grid_constant__ CUtensorMap tm_x, const ConvOutParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* sA = smem;                                              // ring of image rows
  uint8_t* sB = sA + CO_RING * CO_ROW_BYTES;                       // ks x [16 rows x 128 B] weights, swizzled
  float* sP = reinterpret_cast<float*>(sB + 5 * 2048);             // 2 sets x 2 x [16][CO_PPITCH] transpose buffers
  float* s_scale = sP + 4 * 16 * CO_PPITCH;
  float* s_shift = s_scale + 8;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_shift + 8);
  uint64_t* empty = full + CO_RING;
  uint64_t* tfull = empty + CO_RING;
  uint64_t* tempty = tfull + CO_ACC;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + CO_ACC);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int ks = KS_ ? KS_ : p.ks;
  const int cout = COUT_ ? COUT_ : p.cout;
  const int R = ks >> 1;

  // weights: B_ky[n = kx*cout + co][k = ci], K-major, 128-byte swizzle, 16 rows (zero padded)
  for (int i = threadIdx.x; i < ks * 16 * 64; i += CO_THREADS) {
    const int ci = i & 63, n = (i >> 6) & 15, ky = i >> 10;
    const int kx = n / cout, co = n - kx * cout;
    float v = 0.f;
    if (kx < ks) v = p.w[co * p.w_so + ci * p.w_si + ky * p.w_sh + kx * p.w_sw];
    *reinterpret_cast<uint16_t*>(sB + ky * 2048 + n * 128 + (((ci >> 3) ^ (n & 7)) << 4) + (ci & 7) * 2) = cvt16(v, p.f16);
  }
  if (threadIdx.x < 8) {
    float sc = 0.f, sh = 0.f;
    if ((int)threadIdx.x < cout) epi_scale_shift(p.epi, threadIdx.x, sc, sh);
    s_scale[threadIdx.x] = sc;
    s_shift[threadIdx.x] = sh;
  }
  ptx::fence_proxy_async();
  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tm_x);
  if (warp == 1 && lane == 0) {
    // both MMA issuers release a ring row (each commits at every output row, see below)
    for (int s = 0; s < CO_RING; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 2); }
    for (int a = 0; a < CO_ACC; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) { ptx::tmem_alloc(tmem_ptr, 128); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  const int per_img = p.strips * p.row_blocks;
  const int num_work = p.N * per_img;
  const int xstep = CO_MPIX - 2 * R;

  if (warp == 0) {
    // ===================== TMA producer: image rows y0 - R .. y0 + rows + R - 1 of the strip =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
        const int n = wi / per_img, r = wi - n * per_img;
        const int strip = r % p.strips, rb = r / p.strips;
        const int x0 = strip * xstep - R, y0 = rb * CO_RB;
        const int rows = min(CO_RB, p.H - y0);
        for (int i = 0; i < rows + 2 * R; ++i) {
          ptx::mbar_wait(&empty[s], ph ^ 1);
          ptx::mbar_arrive_expect_tx(&full[s], CO_ROW_BYTES);
          ptx::tma_load_4d(sA + s * CO_ROW_BYTES, &tm_x, &full[s], 0, x0, y0 - R + i, n);
          if (++s == CO_RING) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == CO_ISSUER1_WARP) {
    // ===================== MMA issuers =====================
    // An M128 x N16 x K16 MMA occupies the tensor pipe for ~36 clocks (4.5 KB of operands at 128 B/clk) but one warp
    // issues a tcgen05.mma only every ~51 (csrc/experiments/umma_rate_probe.cu), and one set of four epilogue warps
    // needs about as long per output row as the 20 MMAs take.  Output rows ALTERNATE between two (issuer, epilogue set)
    // pairs; each row has its own accumulator stage.  Both issuers walk every row -- waits and ring bookkeeping stay in
    // step -- and both commit the release of the oldest ring row (it was read by MMAs of both; a commit only tracks its
    // own thread's).
    const int issuer = warp == 1 ? 0 : 1;
    const int fmt = p.f16 ? 0 : 1;
    const uint32_t idesc = ptx::umma_idesc_f16(128, 16, fmt, fmt, 0, 0);
    const uint32_t sA_addr = ptx::smem_u32(sA), sB_addr = ptx::smem_u32(sB);
    int head = 0;            // ring slot of the oldest row still needed (input row of tap ky = 0)
    int waited = 0;          // ring slot of the next row to wait for
    uint32_t wph = 0;
    int acc = 0; uint32_t accph = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int n = wi / per_img, r = wi - n * per_img;
      const int rb = r / p.strips;
      const int rows = min(CO_RB, p.H - rb * CO_RB);
      (void)n;
      for (int j = 0; j < rows; ++j) {
        // rows j .. j + 2R of this work item must have landed (the first output row waits for 2R + 1 of them)
        const int need = (j == 0) ? 2 * R + 1 : 1;
        for (int q = 0; q < need; ++q) {
          ptx::mbar_wait(&full[waited], wph);
          if (++waited == CO_RING) { waited = 0; wph ^= 1; }
        }
        if ((j & 1) == issuer) {
          ptx::mbar_wait(&tempty[acc], accph ^ 1);
          ptx::tc_fence_after();
          const uint32_t d0 = tmem_base + (uint32_t)(acc * 16);
#pragma unroll
          for (int ky = 0; ky < (KS_ ? KS_ : 5); ++ky) {
            if (ky >= ks) break;
            int slot = head + ky;
            if (slot >= CO_RING) slot -= CO_RING;
            const uint64_t da = ptx::umma_desc_sw128(sA_addr + (uint32_t)(slot * CO_ROW_BYTES), 1024);
            const uint64_t db = ptx::umma_desc_sw128(sB_addr + (uint32_t)(ky * 2048), 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_elect(d0, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)(ky | k));
          }
          ptx::umma_commit_elect(&tfull[acc]);
        }
        ptx::umma_commit_elect(&empty[head]);          // the oldest row is not used by later output rows
        if (++head == CO_RING) head = 0;
        if (++acc == CO_ACC) { acc = 0; accph ^= 1; }
      }
      // the 2R halo rows at the bottom of the block are still resident: release them
      for (int q = 0; q < 2 * R; ++q) {
        ptx::umma_commit_elect(&empty[head]);
        if (++head == CO_RING) head = 0;
      }
    }
  } else {
    // ===================== epilogue sets (warps 2..5 / 6..9): lane = pixel of the row segment =====================
    const int eset = (warp - 2) >> 2;          // rows j with (j & 1) == eset
    const int quarter = warp & 3;
    const int xl = quarter * 32 + lane;
    int acc = 0, pb = 0; uint32_t accph = 0;
    for (int wi = blockIdx.x; wi < num_work; wi += gridDim.x) {
      const int n = wi / per_img, r = wi - n * per_img;
      const int strip = r % p.strips, rb = r / p.strips;
      const int x0 = strip * xstep - R, y0 = rb * CO_RB;
      const int rows = min(CO_RB, p.H - y0);
      const int X = x0 + xl;
      const bool valid = xl >= R && xl < CO_MPIX - R && X < p.W;
      for (int j = 0; j < rows; ++j) {
        if ((j & 1) != eset) {                  // the other set's row
          if (++acc == CO_ACC) { acc = 0; accph ^= 1; }
          continue;
        }
        ptx::mbar_wait(&tfull[acc], accph);
        ptx::tc_fence_after();
        uint32_t v[16];
        ptx::tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 16), v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
        if (++acc == CO_ACC) { acc = 0; accph ^= 1; }
        float* P = sP + (eset * 2 + pb) * 16 * CO_PPITCH;
#pragma unroll
        for (int c = 0; c < 16; ++c) P[c * CO_PPITCH + xl] = __uint_as_float(v[c]);
        ptx::named_bar_sync(1 + eset, 128);
        if (valid) {
          const int Y = y0 + j;
#pragma unroll
          for (int co = 0; co < (COUT_ ? COUT_ : 8); ++co) {
            if (co >= cout) break;
            float s = 0.f;
#pragma unroll
            for (int kx = 0; kx < (KS_ ? KS_ : 5); ++kx)
              if (kx < ks) s += P[(kx * cout + co) * CO_PPITCH + xl + kx - R];
            float o = apply_act(fmaf(s, s_scale[co], s_shift[co]), p.epi.act);
            if (p.epi.res.ptr) o += p.epi.res.ld(p.epi.res.off(n, co, Y, X));
            if (OUT_ == OFA_F32) reinterpret_cast<float*>(p.y.ptr)[p.y.off(n, co, Y, X)] = o;
            else if (OUT_ == OFA_U8)
              reinterpret_cast<uint8_t*>(p.y.ptr)[p.y.off(n, co, Y, X)] = (uint8_t)rintf(fminf(fmaxf(o, 0.f), 1.f) * 255.f);
            else p.y.st(p.y.off(n, co, Y, X), o);
          }
        }
        pb ^= 1;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 128); }
}

// ==================================================================================================
// (2) stem: cin <= 4 -> 64, CUDA cores
// ==================================================================================================
constexpr int ST_TW = 64, ST_TH = 16;     // output tile (pixels)
constexpr int ST_THREADS = 256;           // 16 x-groups of 4 pixels, 4 row groups, 4 channel groups of 16
constexpr int ST_MAXK = 5, ST_MAXC = 4;

struct StemParams {
  TV x, y;
  const float* w; long long w_so, w_si, w_sh, w_sw;
  int cin, ks, f16, store;
  Epi epi;
  int tiles_x, tiles_y, num_tiles;
};

// CGS = channel groups of 16: 4 -> 64 output channels (the stems), 1 -> 16 (X4's first conv, stored pixel-unshuffled)
template <int CGS>
__global__ void __launch_bounds__(ST_THREADS)
conv_stem_kernel(const StemParams p) {
  __shared__ float s_in[ST_MAXC][ST_TH + ST_MAXK - 1][ST_TW + ST_MAXK - 1 + 1];
  __shared__ __align__(16) float s_w[ST_MAXK * ST_MAXK * ST_MAXC][64];   // [ky][kx][ci][co]
  __shared__ float s_scale[64], s_shift[64];
  const int tid = threadIdx.x;
  const int ks = p.ks, R = ks >> 1, cin = p.cin;
  constexpr int COUT = 16 * CGS;
  const int per_img = p.tiles_x * p.tiles_y;
  const int H = p.x.h, W = p.x.w;

  // weights once per (persistent) block, read in memory order (coalesced for a dense parameter)
  for (int i = tid; i < ks * ks * cin * COUT; i += ST_THREADS) {
    const int kx = i % ks;
    int r = i / ks;
    const int ky = r % ks; r /= ks;
    const int ci = r % cin, co = r / cin;
    // channel co = cg*16 + k*4 + e is stored at float4 slot k*4 + cg: the four channel groups of a quarter warp then
    // read four CONSECUTIVE 16-byte slots (no bank conflict; [co] order put cg 0/2 and 1/3 on the same banks)
    const int slot = ((co & 15) >> 2) * CGS + (co >> 4);
    s_w[(ky * ks + kx) * cin + ci][slot * 4 + (co & 3)] = p.w[co * p.w_so + ci * p.w_si + ky * p.w_sh + kx * p.w_sw];
  }
  if (tid < COUT) {
    float sc, sh;
    epi_scale_shift(p.epi, tid, sc, sh);
    s_scale[tid] = sc; s_shift[tid] = sh;
  }
  const int hh = ST_TH + ks - 1, hw = ST_TW + ks - 1;
  for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
  const int n = t / per_img, r = t - n * per_img;
  const int y0 = (r / p.tiles_x) * ST_TH, x0 = (r % p.tiles_x) * ST_TW;
  __syncthreads();                                  // previous tile's readers are done with s_in (and s_w is written)
  for (int i = tid; i < cin * hh * hw; i += ST_THREADS) {
    const int xx = i % hw, q = i / hw;
    const int yy = q % hh, ci = q / hh;
    const int iy = y0 + yy - R, ix = x0 + xx - R;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = p.x.ld(p.x.off(n, ci, iy, ix));
    s_in[ci][yy][xx] = v;
  }
  __syncthreads();

  // thread -> 4 consecutive pixels of one row x 16 channels; 4 rows per thread in turn
  const int cg = tid % CGS, xg = (tid / CGS) & 15, rg = tid / (CGS * 16);
  for (int ry = rg; ry < ST_TH; ry += ST_THREADS / (CGS * 16)) {
    float2 acc[4][8];            // channel pairs: packed fp32x2 FMAs (fma.rn.f32x2) double the CUDA-core rate
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] = make_float2(0.f, 0.f);
    for (int ky = 0; ky < ks; ++ky) {
      for (int ci = 0; ci < cin; ++ci) {
        float in[4 + ST_MAXK - 1];
#pragma unroll
        for (int i = 0; i < 4 + ST_MAXK - 1; ++i) in[i] = (i < 4 + ks - 1) ? s_in[ci][ry + ky][xg * 4 + i] : 0.f;
#pragma unroll
        for (int kx = 0; kx < ST_MAXK; ++kx) {
          if (kx < ks) {
            const float4* wp = reinterpret_cast<const float4*>(&s_w[(ky * ks + kx) * cin + ci][0]);
            float2 wv[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 f = wp[q * CGS + cg]; wv[2 * q] = make_float2(f.x, f.y); wv[2 * q + 1] = make_float2(f.z, f.w); }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 xv = make_float2(in[i + kx], in[i + kx]);
#pragma unroll
              for (int c = 0; c < 8; ++c) ptx::ffma2(acc[i][c], xv, wv[c]);
            }
          }
        }
      }
    }
    const int Y = y0 + ry;
    if (Y < H) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int X = x0 + xg * 4 + i;
        if (X < W) {
          float o[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int co = cg * 16 + c;
            const float av = (c & 1) ? acc[i][c >> 1].y : acc[i][c >> 1].x;
            o[c] = apply_act(fmaf(av, s_scale[co], s_shift[co]), p.epi.act);
            if (p.store == OFA_STORE_PLAIN && p.epi.res.ptr) o[c] += p.epi.res.ld(p.epi.res.off(n, co, Y, X));
          }
          if (p.store != OFA_STORE_PLAIN) {                   // PixelUnshuffle / PixelShuffle store (layers.py:94-98)
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              int oc, oh, ow;
              store_coord(p.store, cg * 16 + c, Y, X, oc, oh, ow);
              float v = o[c];
              if (p.epi.res.ptr) v += p.epi.res.ld(p.epi.res.off(n, oc, oh, ow));
              p.y.st(p.y.off(n, oc, oh, ow), v);
            }
          } else if (p.y.dtype != OFA_F32 && p.y.sc == 1) {   // NHWC 16-bit: two 16-byte stores per pixel
            uint32_t pk[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) pk[q] = pack16(o[2 * q], o[2 * q + 1], p.f16);
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.y.ptr) + p.y.off(n, cg * 16, Y, X));
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) p.y.st(p.y.off(n, cg * 16 + c, Y, X), o[c]);
          }
        }
      }
    }
  }
  }   // tile loop
}

}  // namespace

bool conv_out_rows_supported(const OfaConvArgs* a) {
  if (a->flip || a->store != OFA_STORE_PLAIN) return false;
  if (!is_16bit(a->x.dtype) || !is_nhwc_dense(&a->x) || a->cin != 64) return false;
  if (a->ks != 3 && a->ks != 5) return false;
  if (a->cout < 1 || a->cout > 8 || a->ks * a->cout > 16) return false;
  if (!a->w || (reinterpret_cast<uintptr_t>(a->x.ptr) & 15)) return false;
  if (a->x.n <= 0 || a->x.h <= 0 || a->x.w <= 0) return false;
  return true;
}

int launch_conv_out_rows(const OfaConvArgs* a, cudaStream_t st) {
  ConvOutParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->x.n; p.H = a->x.h; p.W = a->x.w; p.cout = a->cout; p.ks = a->ks;
  p.f16 = a->x.dtype == OFA_F16 ? 1 : 0;
  p.w = a->w; p.w_so = a->w_so; p.w_si = a->w_si; p.w_sh = a->w_sh; p.w_sw = a->w_sw;
  const int R = a->ks / 2;
  p.strips = (p.W + (CO_MPIX - 2 * R) - 1) / (CO_MPIX - 2 * R);
  p.row_blocks = (p.H + CO_RB - 1) / CO_RB;
  p.epi = make_epi(&a->epi);
  p.y = make_tv(&a->y);
  CUtensorMap tx;
  uint64_t dims[4] = {64, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
  uint64_t strides[3] = {128, (uint64_t)p.W * 128, (uint64_t)p.H * p.W * 128};
  uint32_t box[4] = {64, CO_MPIX, 1, 1};
  int rc = encode_tmap(&tx, p.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->x.ptr, dims,
                       strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const size_t smem = 1024 + CO_RING * CO_ROW_BYTES + 5 * 2048 + 4 * 16 * CO_PPITCH * 4 + 64 + (2 * CO_RING + 2 * CO_ACC) * 8 + 64;
  const int num_work = p.N * p.strips * p.row_blocks;
  int grid = sm_count();
  if (grid > num_work) grid = num_work;
#define OFA_CO_LAUNCH(K_, C_, O_)                                                                                    \
  do {                                                                                                              \
    OFA_CUDA(cudaFuncSetAttribute(conv_out_rows_kernel<K_, C_, O_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                  (int)smem));                                                                      \
    conv_out_rows_kernel<K_, C_, O_><<<grid, CO_THREADS, smem, st>>>(tx, p);                                        \
  } while (0)
  const int od = a->y.dtype;
  if (p.ks == 5 && p.cout == 3 && od == OFA_F32) OFA_CO_LAUNCH(5, 3, OFA_F32);
  else if (p.ks == 5 && p.cout == 3 && od == OFA_U8) OFA_CO_LAUNCH(5, 3, OFA_U8);
  else if (p.ks == 3 && p.cout == 3 && od == OFA_F32) OFA_CO_LAUNCH(3, 3, OFA_F32);
  else if (p.ks == 3 && p.cout == 3 && od == OFA_U8) OFA_CO_LAUNCH(3, 3, OFA_U8);
  else OFA_CO_LAUNCH(0, 0, -1);
#undef OFA_CO_LAUNCH
  return check_launch("conv_out_rows_kernel");
}

bool conv_stem_supported(const OfaConvArgs* a) {
  if (a->flip) return false;
  if (a->cin < 1 || a->cin > ST_MAXC || (a->cout != 64 && a->cout != 16)) return false;
  if (a->store == OFA_STORE_PIXELSHUFFLE2) return false;
  if (a->store == OFA_STORE_PLAIN && a->cout != 64 && a->y.dtype != OFA_F32 && a->y.sc == 1) return false;
  if (a->ks != 3 && a->ks != 5) return false;
  if (!a->w || a->x.n <= 0 || a->x.h <= 0 || a->x.w <= 0) return false;
  if (a->y.dtype != OFA_F32 && a->y.sc == 1 && ((reinterpret_cast<uintptr_t>(a->y.ptr) & 15) || a->y.sw % 8 || a->y.sh % 8 || a->y.sn % 8))
    return false;
  return true;
}

int launch_conv_stem(const OfaConvArgs* a, cudaStream_t st) {
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.x = make_tv(&a->x); p.y = make_tv(&a->y);
  p.w = a->w; p.w_so = a->w_so; p.w_si = a->w_si; p.w_sh = a->w_sh; p.w_sw = a->w_sw;
  p.cin = a->cin; p.ks = a->ks; p.f16 = a->y.dtype == OFA_F16 ? 1 : 0; p.store = a->store;
  p.epi = make_epi(&a->epi);
  p.tiles_x = (a->x.w + ST_TW - 1) / ST_TW;
  p.tiles_y = (a->x.h + ST_TH - 1) / ST_TH;
  const long long tiles = (long long)a->x.n * p.tiles_x * p.tiles_y;
  if (tiles >= (1ll << 31)) return fail(OFA_ERR_UNSUPPORTED, "conv_stem: too many tiles");
  p.num_tiles = (int)tiles;
  long long blocks = 4LL * sm_count();            // persistent: the 19 KB weight tensor is loaded once per block
  if (blocks > tiles) blocks = tiles;
  if (a->cout == 64) conv_stem_kernel<4><<<(unsigned)blocks, ST_THREADS, 0, st>>>(p);
  else conv_stem_kernel<1><<<(unsigned)blocks, ST_THREADS, 0, st>>>(p);
  return check_launch("conv_stem_kernel");
}

}  // namespace ofa
